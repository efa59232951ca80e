"""Pin the CPU oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU-only; runs in the `-m "not gpu"` suite."""
import json
import random

import numpy as np
import pytest
import torch

import oracle
from oracle import pd as PD
from oracle.filters import ZFilter, RewardFilter
from oracle import nets
from oracle.gae import gae_from_values, gae_reference_fp64
from oracle.ppo import OraclePPOLearner
from oracle.ppo_rnn import OraclePPOLearnerRNN
from oracle.ppo_pixel import OraclePPOLearnerPixel
from oracle.ddpg import OracleDDPGLearner
from oracle.replay import FIFO, Uniform, MT19937
from oracle.windowing import multistep_windows, ssar_nstep
from oracle.aggregator import multistep_aggregate, ssar_aggregate
from oracle.agent import ppo_act, ddpg_act, ppo_act_rnn, ddpg_act_ou

torch.set_num_threads(1)
T = torch.tensor


def test_diag_gauss(golden):
    g = golden('pd')
    a, p0, p1 = T(g['a']), T(g['p0']), T(g['p1'])
    d = a.shape[1]
    np.testing.assert_array_equal(PD.loglikelihood(a, p0, d).numpy(), g['loglik'])
    np.testing.assert_array_equal(PD.likelihood(a, p0, d).numpy(), g['lik'])
    assert (g['lik'][:3] == np.float32(1e-5)).all()          # the clamp engaged in the fixture
    np.testing.assert_array_equal(PD.kl(p0, p1, d).numpy(), g['kl01'])
    np.testing.assert_array_equal(PD.entropy(p0, d).numpy(), g['ent'])


def test_zfilter(golden):
    g = golden('zfilter')
    zf = ZFilter(g['x0'].shape[1])
    x0 = T(g['x0'])
    np.testing.assert_array_equal(zf.forward(x0).numpy(), g['y_init'])
    for i, x in enumerate(g['xs']):
        zf.update(T(x))
        st = np.concatenate([zf.running_sum.numpy(), zf.running_sumsq.numpy(), zf.count.numpy()])
        np.testing.assert_array_equal(st, g['states'][i])
        np.testing.assert_array_equal(zf.forward(x0).numpy(), g['outs'][i])
    np.testing.assert_array_equal(zf.running_mean(), g['running_mean'])
    np.testing.assert_array_equal(zf.running_std(), g['running_std'])
    np.testing.assert_array_equal(zf.running_square(), g['running_square'])


def test_reward_filter_overwrite_quirk(golden):
    g = golden('rfilter')
    rf = RewardFilter()
    for i, x in enumerate(g['r']):
        np.testing.assert_array_equal(rf.forward(T(x)).numpy(), g['outs'][i])
        rf.update(T(x))
        np.testing.assert_allclose([rf.count.item(), rf.running_sum.item(), rf.running_sumsq.item()],
                                   g['states'][i], rtol=0, atol=0)
    assert rf.reward_mean() == float(g['reward_mean'])


def _ppo_model(sd):
    actor = nets.params_from_state(sd, 'actor/model/')
    critic = nets.params_from_state(sd, 'critic/model/')
    log_var = T(sd['actor/log_var'])
    zf = None
    if 'z_filter/count' in sd:
        zf = ZFilter(len(sd['z_filter/running_sum'])).load(sd['z_filter/running_sum'], sd['z_filter/running_sumsq'],
                                                           sd['z_filter/count'])
    return actor, log_var, critic, zf


@pytest.mark.parametrize('tag', ['mlp', 'mlp_nonorm'])
def test_gae_mlp(golden, tag):
    g = golden('gae_' + tag)
    actor, log_var, critic, zf = _ppo_model(g.sub('model/'))
    obs = T(g['obs'])
    cat = torch.cat([obs, T(g['obs_next'])], 1)
    B, n1, D = cat.shape
    with torch.no_grad():
        values = nets.ppo_critic(zf.forward(cat.view(-1, D)), critic).view(B, n1)
    np.testing.assert_array_equal(values.numpy(), g['values_raw'])
    rewards = T(g['rewards'], dtype=torch.float32)
    adv, ret = gae_from_values(rewards, values, T(g['dones']), float(g['gamma']), float(g['lam']),
                               norm_adv=bool(g['norm_adv']))
    np.testing.assert_array_equal(adv.numpy(), g['adv'])
    np.testing.assert_array_equal(ret.numpy(), g['ret'])
    if not bool(g['norm_adv']):
        a64, r64 = gae_reference_fp64(rewards, values, T(g['dones']), float(g['gamma']), float(g['lam']))
        scale = max(1.0, float(a64.pow(2).mean().sqrt()))
        assert (adv.view(-1).double() - a64).abs().max() <= 1e-5 * scale
        assert (ret.view(-1).double() - r64).abs().max() <= 1e-5 * max(1.0, float(r64.pow(2).mean().sqrt()))


def test_gae_rnn_mode_given_values(golden):
    """RNN branch (ppo.py:389-406): the oracle is fed the reference's critic values (the LSTM stem is a
    'next' row) and must reproduce the horizon-windowed advantages / returns."""
    g = golden('gae_rnn')
    adv, ret = gae_from_values(T(g['rewards'], dtype=torch.float32), T(g['values_raw']), T(g['dones']),
                               float(g['gamma']), float(g['lam']), horizon=int(g['horizon']),
                               norm_adv=bool(g['norm_adv']))
    np.testing.assert_array_equal(adv.numpy(), g['adv'])
    np.testing.assert_array_equal(ret.numpy(), g['ret'])


def _state_matches(L, after, atol=0.0):
    got = {}
    for i, (w, b) in enumerate(L.actor):
        got['actor/model/seq/%d/weight' % (2 * i)] = w
        got['actor/model/seq/%d/bias' % (2 * i)] = b
    for i, (w, b) in enumerate(L.critic):
        got['critic/model/seq/%d/weight' % (2 * i)] = w
        got['critic/model/seq/%d/bias' % (2 * i)] = b
    got['actor/log_var'] = L.log_var
    got['z_filter/running_sum'] = L.zf.running_sum
    got['z_filter/running_sumsq'] = L.zf.running_sumsq
    got['z_filter/count'] = L.zf.count
    for k, v in got.items():
        np.testing.assert_allclose(v.detach().numpy(), after[k], rtol=0, atol=atol, err_msg=k)


@pytest.mark.parametrize('tag', ['clip', 'adapt', 'clip_biglr', 'adapt_biglr', 'clip_rfilter'])
def test_ppo_learn(golden, tag):
    g = golden('ppo_learn_' + tag)
    cfg = g.js('cfg')
    hyper = g.js('hyper')
    stats = g.js('stats')
    actor, log_var, critic, zf = _ppo_model(g.sub('init/'))
    L = OraclePPOLearner(actor, log_var, critic, zf, cfg['A'], cfg['n_step'], cfg['B'], ppo_mode=cfg['mode'],
                         lr_actor=cfg['lr'], lr_critic=cfg['lr'], exp_interval=cfg['exp_interval'],
                         use_r_filter=cfg['use_r_filter'], reward_scale=cfg['reward_scale'])
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        st = L.learn(dict(obs=b['obs'], obs_next=b['obs_next'], actions=b['actions'], rewards=b['rewards'],
                          dones=b['dones'], pd=b['pd']))
        L.publish_parameter()
        assert L.n_policy_epochs[-1] == hyper[it]['n_policy_epochs']
        for k, v in stats[it].items():
            assert st[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        _state_matches(L, g.sub('it%d/after/' % it), atol=1e-7)
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)
        assert L.exp_counter == hyper[it]['exp_counter']
        np.testing.assert_allclose(L.ref_log_var.numpy(), b['ref/actor/log_var'], atol=1e-7)
        np.testing.assert_allclose(L.ref_zf.count.numpy(), b['ref/z_filter/count'], atol=0)
    assert any(h['n_policy_epochs'] < cfg['epoch_policy'] for h in hyper) == (tag == 'clip_biglr')


@pytest.mark.parametrize('tag', ['rnn_clip', 'rnn_adapt', 'rnn_adapt_biglr'])
def test_ppo_learn_rnn_mode(golden, tag):
    """RNN mode (the reference's default PPO config): LSTM stem trained by both optimisers, horizon GAE over
    eff_len positions, initial cells from onetime_infos -- learn() + publish against goldens produced by running the
    real reference (tests/golden/make_golden.py::gen_ppo_learn_rnn).  Pins the oracle for SURVEY §8(f) rank 2; the
    CUDA path for this mode is not built yet."""
    g = golden('ppo_learn_' + tag)
    cfg, hyper, stats = g.js('cfg'), g.js('hyper'), g.js('stats')
    init = g.sub('init/')
    actor, log_var, critic, zf = _ppo_model(init)
    lstm = {k.split('/', 1)[1]: init[k] for k in init.keys() if k.startswith('rnn_stem/')}
    L = OraclePPOLearnerRNN(actor, log_var, critic, zf, lstm, cfg['A'], cfg['n_step'], cfg['B'], cfg['horizon'],
                            cfg['rnn_hidden'], cfg['rnn_layer'], ppo_mode=cfg['mode'], lr_actor=cfg['lr'],
                            lr_critic=cfg['lr'], exp_interval=cfg['exp_interval'])
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        st = L.learn(dict(obs=b['obs'], obs_next=b['obs_next'], actions=b['actions'], rewards=b['rewards'],
                          dones=b['dones'], pd=b['pd'], h0=b['h0'], c0=b['c0']))
        np.testing.assert_allclose(L.last_adv.numpy(), b['adv'], rtol=0, atol=1e-6)
        np.testing.assert_allclose(L.last_ret.numpy(), b['ret'], rtol=0, atol=1e-6)
        assert L.last_adv.shape == (cfg['B'], cfg['n_step'] - cfg['horizon'] + 1)
        L.publish_parameter()
        assert L.n_policy_epochs[-1] == hyper[it]['n_policy_epochs']
        for k, v in stats[it].items():
            assert st[k] == pytest.approx(v, rel=1e-5, abs=1e-6), k
        _state_matches(L, g.sub('it%d/after/' % it), atol=1e-6)
        after = g.sub('it%d/after/' % it)
        for name, p in L.rnn.named_parameters():
            np.testing.assert_allclose(p.detach().numpy(), after['rnn_stem/' + name], rtol=0, atol=1e-6, err_msg=name)
        for name, p in L.ref_rnn.named_parameters():
            np.testing.assert_allclose(p.detach().numpy(), b['ref/rnn_stem/' + name], rtol=0, atol=1e-6, err_msg=name)
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)
        assert L.exp_counter == hyper[it]['exp_counter']


@pytest.mark.parametrize('tag', ['pixel_clip', 'pixel_adapt'])
def test_ppo_learn_pixel_mode(golden, tag):
    """Pixel mode (BASELINE cfg 4 shape, scaled down): uint8 frames, CNN stem shared by actor and critic and trained
    by both optimisers, no z-filter -- against goldens produced by the real reference.  Pins the oracle for the CNN
    branch of SURVEY §8 rows a2/a26; the CUDA conv kernels are not built yet."""
    g = golden('ppo_learn_' + tag)
    cfg, hyper, stats = g.js('cfg'), g.js('hyper'), g.js('stats')
    init = g.sub('init/')
    actor = nets.params_from_state(init, 'actor/model/')
    critic = nets.params_from_state(init, 'critic/model/')
    log_var = T(init['actor/log_var'])
    conv = [(T(init['cnn_stem/model/seq/%d/weight' % i]), T(init['cnn_stem/model/seq/%d/bias' % i])) for i in (0, 2)]
    fc = (T(init['cnn_stem/model/seq/5/weight']), T(init['cnn_stem/model/seq/5/bias']))
    L = OraclePPOLearnerPixel(actor, log_var, critic, conv, [4, 2], fc, cfg['A'], cfg['n_step'], cfg['B'],
                              ppo_mode=cfg['mode'], lr_actor=cfg['lr'], lr_critic=cfg['lr'],
                              exp_interval=cfg['exp_interval'])
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        assert b['obs'].dtype == np.uint8
        st = L.learn(dict(obs=b['obs'], obs_next=b['obs_next'], actions=b['actions'], rewards=b['rewards'],
                          dones=b['dones'], pd=b['pd']))
        np.testing.assert_allclose(L.last_adv.numpy(), b['adv'], rtol=0, atol=2e-6)
        np.testing.assert_allclose(L.last_ret.numpy(), b['ret'], rtol=0, atol=2e-6)
        L.publish_parameter()
        for k, v in stats[it].items():
            assert st[k] == pytest.approx(v, rel=1e-5, abs=1e-6), k
        after = g.sub('it%d/after/' % it)
        for i, (w, bb) in zip((0, 2), L.conv):
            np.testing.assert_allclose(w.detach().numpy(), after['cnn_stem/model/seq/%d/weight' % i], rtol=0, atol=1e-6)
            np.testing.assert_allclose(bb.detach().numpy(), after['cnn_stem/model/seq/%d/bias' % i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(L.fc[0].detach().numpy(), after['cnn_stem/model/seq/5/weight'], rtol=0, atol=1e-6)
        for i, (w, bb) in enumerate(L.actor):
            np.testing.assert_allclose(w.detach().numpy(), after['actor/model/seq/%d/weight' % (2 * i)], rtol=0, atol=1e-6)
        for i, (w, bb) in enumerate(L.critic):
            np.testing.assert_allclose(w.detach().numpy(), after['critic/model/seq/%d/weight' % (2 * i)], rtol=0, atol=1e-6)
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)


def test_ddpg_act_ou_noise(golden):
    """Ornstein-Uhlenbeck exploration (action_noise.py:22-39): float64 state carried between steps, reset in
    pre_episode()."""
    g = golden('ddpg_act_ou')
    actor = nets.params_from_state(g.sub('model/'), 'actor/model/')
    x = np.zeros(3)
    for i in range(len(g['obs'])):
        if i == int(g['reset_at']):
            x = np.zeros(3)
        a, x = ddpg_act_ou(g['obs'][i], actor, x, float(g['sigma']), float(g['theta']), float(g['dt']),
                           unit_noise=g['unit_noise'][i])
        np.testing.assert_array_equal(x, g['ou_states'][i])
        np.testing.assert_array_equal(a, g['actions'][i])
    assert np.abs(g['ou_states'][4]).max() > np.abs(g['ou_states'][5]).max() * 0.5   # the walk restarted at reset_at


def test_ppo_act_rnn_mode(golden):
    """Actor side of RNN mode: cells carried between steps, the PRE-step cells are what travels as onetime_infos,
    reset() zeroes them (ppo_agent.py:84-93,133-137,169-183)."""
    g = golden('ppo_act_rnn')
    sd = g.sub('model/')
    actor, log_var, _, zf = _ppo_model(sd)
    Hd = int(g['rnn_hidden'])
    lstm = torch.nn.LSTM(11, Hd, 1, batch_first=True)
    lstm.load_state_dict({k.split('/', 1)[1]: T(sd[k]) for k in sd.keys() if k.startswith('rnn_stem/')})
    zero = lambda: (torch.zeros(1, 1, Hd), torch.zeros(1, 1, Hd))   # noqa: E731
    cells = zero()
    for i in range(len(g['obs'])):
        if i == int(g['reset_at']):
            cells = zero()
        a, pdv, onetime, cells = ppo_act_rnn(g['obs'][i], actor, log_var, zf, lstm, cells, float(g['noise']), eps=g['eps'][i])
        # torch's CPU LSTM differs by 1 ulp between thread counts (golden: default threads, here: 1)
        np.testing.assert_allclose(onetime[0], g['h_before'][i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(onetime[1], g['c_before'][i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(pdv, g['pds'][i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(a, g['actions'][i], rtol=0, atol=1e-7)
    assert np.abs(g['h_before'][1]).max() > 0 and np.abs(g['h_before'][int(g['reset_at'])]).max() == 0


def _ddpg_nets(sd):
    actor = nets.params_from_state(sd, 'actor/model/')
    c0 = nets.params_from_state(sd, 'critic/model_obs/', 1)
    c12 = nets.params_from_state(sd, 'critic/model_concat/', 2)
    return actor, c0 + c12


@pytest.mark.parametrize('tag', ['hard', 'soft_clipcritic'])
def test_ddpg_optimize(golden, tag):
    g = golden('ddpg_optimize_' + tag)
    cfg = g.js('cfg')
    a, c = _ddpg_nets(g.sub('init/model/'))
    at, ct = _ddpg_nets(g.sub('init/target/'))
    L = OracleDDPGLearner(a, c, at, ct, gamma=cfg['gamma'], n_step=cfg['n_step'], lr_actor=cfg['lr_actor'],
                          lr_critic=cfg['lr_critic'], clip_actor=cfg['clip_actor'], actor_clip=cfg['actor_clip'],
                          clip_critic=cfg['clip_critic'], critic_clip=cfg['critic_clip'],
                          target_type=cfg['target']['type'], target_interval=cfg['target'].get('interval', 0),
                          tau=cfg['target'].get('tau', 0.0))
    stats = g.js('stats')
    for it in range(3):
        b = g.sub('it%d/' % it)
        st = L.optimize(b['obs'], b['actions'], b['rewards'], b['obs_next'], b['dones'])
        for k, v in stats[it].items():
            assert st[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        ea, ec = _ddpg_nets(g.sub('it%d/model/' % it))
        eat, ect = _ddpg_nets(g.sub('it%d/target/' % it))
        for got, exp in [(L.actor, ea), (L.critic, ec), (L.actor_t, eat), (L.critic_t, ect)]:
            for (w, bb), (we, be) in zip(got, exp):
                np.testing.assert_allclose(w.detach().numpy(), we.numpy(), rtol=0, atol=1e-7)
                np.testing.assert_allclose(bb.detach().numpy(), be.numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize('tag', ['td3_double', 'td3_double_reg'])
def test_ddpg_optimize_td3_options(golden, tag):
    """TD3 options (ddpg.py:267-283,298-321): second critic + its target, y = min(y, y2); the smoothing noise is added
    to the target action only AFTER Q'_1 was evaluated (a reference quirk the oracle keeps)."""
    g = golden('ddpg_optimize_' + tag)
    cfg = g.js('cfg')
    a, c = _ddpg_nets(g.sub('init/model/'))
    at, ct = _ddpg_nets(g.sub('init/target/'))
    crit = lambda sd: (nets.params_from_state(sd, 'critic/model_obs/', 1) +          # noqa: E731  (model2 holds no actor)
                       nets.params_from_state(sd, 'critic/model_concat/', 2))
    c2, c2t = crit(g.sub('init/model2/')), crit(g.sub('init/target2/'))
    L = OracleDDPGLearner(a, c, at, ct, gamma=cfg['gamma'], n_step=cfg['n_step'], lr_actor=cfg['lr_actor'],
                          lr_critic=cfg['lr_critic'], clip_actor=cfg['clip_actor'], actor_clip=cfg['actor_clip'],
                          clip_critic=cfg['clip_critic'], critic_clip=cfg['critic_clip'],
                          target_type=cfg['target']['type'], target_interval=cfg['target'].get('interval', 0),
                          tau=cfg['target'].get('tau', 0.0), critic2=c2, critic2_t=c2t)
    stats = g.js('stats')
    for it in range(3):
        b = g.sub('it%d/' % it)
        noise = b['policy_noise_unclipped'] if cfg['action_regularization'] else None
        st = L.optimize(b['obs'], b['actions'], b['rewards'], b['obs_next'], b['dones'], policy_noise=noise)
        for k, v in stats[it].items():
            assert st[k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        assert 'Q_policy2' in st
        ea, ec = _ddpg_nets(g.sub('it%d/model/' % it))
        ec2, ec2t = crit(g.sub('it%d/model2/' % it)), crit(g.sub('it%d/target2/' % it))
        for got, exp in [(L.actor, ea), (L.critic, ec), (L.critic2, ec2), (L.critic2_t, ec2t)]:
            for (w, bb), (we, be) in zip(got, exp):
                np.testing.assert_allclose(w.detach().numpy(), we.numpy(), rtol=0, atol=1e-7)
                np.testing.assert_allclose(bb.detach().numpy(), be.numpy(), rtol=0, atol=1e-7)


def test_fifo_replay_trace(golden):
    f = golden('replay').js('fifo')
    R = FIFO(f['memory_size'], f['batch_size'])
    nxt = 0
    for (op, k), tr in zip(f['script'], f['trace']):
        if op == 'insert':
            for _ in range(k):
                R.insert(nxt)
                nxt += 1
            assert [len(R), int(R.ready())] == tr
        else:
            assert R.sample(k) == tr


def test_uniform_replay_trace(golden):
    u = golden('replay').js('uniform')
    rng = random.Random(u['seed'])
    mt = MT19937(u['seed'])
    R = Uniform(u['memory_size'], u['sampling_start_size'], rng)
    R2 = Uniform(u['memory_size'], u['sampling_start_size'], mt)
    nxt = 0
    for (op, k), tr in zip(u['script'], u['trace']):
        if op == 'insert':
            for _ in range(k):
                R.insert(nxt)
                R2.insert(nxt)
                nxt += 1
            assert [len(R), int(R.ready()), R.next_idx] == tr
        else:
            assert R.sample(k) == tr
            assert R2.sample(k) == tr


def test_randint_streams(golden):
    s = golden('replay').js('streams')
    for m, exp in s.items():
        if m.startswith('bigseed'):
            mt, rr, mm = MT19937(12345678901234567890), random.Random(12345678901234567890), 1000
        else:
            mt, rr, mm = MT19937(5), random.Random(5), int(m)
        assert [mt.randint(0, mm - 1) for _ in range(len(exp))] == exp
        assert [rr.randint(0, mm - 1) for _ in range(len(exp))] == exp


def test_multistep_windows(golden):
    for case in golden('window_multistep').js('cases'):
        got = multistep_windows(case['ep_lens'], case['n_step'], case['stride'])
        exp = case['windows']
        assert len(got) == len(exp)
        for (obs_ids, nxt, dones), w in zip(got, exp):
            assert obs_ids == w['obs'] and nxt == w['obs_next'] and dones == w['dones']
            assert [float(i) for i in obs_ids] == w['actions'] == w['pd0']
            assert [0.25 * (i + 1) - 3.0 for i in obs_ids] == w['rewards']


def test_ssar_nstep(golden):
    for case in golden('window_ssar').js('cases'):
        got = ssar_nstep(case['ep_lens'], case['n_step'], case['gamma'], lambda g: 0.25 * g - 3.0)
        exp = case['records']
        assert len(got) == len(exp)
        for (o, on, a, r, d), e in zip(got, exp):
            assert (o, on, float(a), d) == (e['obs'], e['obs_next'], e['action'], e['done'])
            assert r == e['reward']            # float64 accumulation order is identical


def test_aggregators(golden):
    g = golden('aggregate')
    dt = g.js('dtypes')
    B = g['ms_in_obs'].shape[0]
    wins = [dict(obs=list(g['ms_in_obs'][b]), obs_next=g['ms_in_obs_next'][b], actions=list(g['ms_in_actions'][b]),
                 rewards=[float(x) for x in g['ms_in_rewards'][b]], dones=[bool(x) for x in g['ms_in_dones'][b]],
                 pd=list(g['ms_in_pd'][b])) for b in range(B)]
    out = multistep_aggregate(wins)
    for k in ['obs', 'obs_next', 'actions', 'rewards', 'dones', 'pd']:
        np.testing.assert_array_equal(out[k], g['ms_' + k])
        assert str(out[k].dtype) == dt['ms_' + k]
    exps = [dict(obs=g['ss_in_obs'][b], obs_next=g['ss_in_obs_next'][b], action=g['ss_in_action'][b],
                 reward=float(g['ss_in_reward'][b]), done=bool(g['ss_in_done'][b])) for b in range(4)]
    out = ssar_aggregate(exps)
    for k in ['obs', 'obs_next', 'actions', 'rewards', 'dones']:
        np.testing.assert_array_equal(out[k], g['ss_' + k])
        assert str(out[k].dtype) == dt['ss_' + k]


def test_ppo_act(golden):
    g = golden('ppo_act')
    actor, log_var, _, zf = _ppo_model(g.sub('model/'))
    for i in range(len(g['obs'])):
        a, pdv = ppo_act(g['obs'][i], actor, log_var, zf, float(g['noise']), eps=g['eps'][i])
        np.testing.assert_array_equal(a, g['actions'][i])
        np.testing.assert_array_equal(pdv, g['pds'][i])
        assert a.dtype == np.float64
        ad, _ = ppo_act(g['obs'][i], actor, log_var, zf, 0.0, deterministic=True)
        np.testing.assert_array_equal(ad, g['actions_det'][i])


def test_ddpg_act(golden):
    g = golden('ddpg_act')
    actor, _ = _ddpg_nets(g.sub('model/'))
    for i in range(len(g['obs'])):
        a = ddpg_act(g['obs'][i], actor, float(g['sigma']), g['unit_noise'][i])
        np.testing.assert_array_equal(a, g['actions'][i])
        assert a.dtype == np.float32
