"""RNN mode on the GPU (the reference's DEFAULT PPO config): LSTM stem kernels vs torch.nn.LSTM, the
RNN-mode PPOLearner against goldens produced by the REFERENCE's PPOLearner (LSTM trained by both optimisers, horizon GAE
over eff_len positions, initial cells from onetime_infos; ppo.py:389-406,507-525), PPOAgent.act with cell hand-off against
the reference golden, and the whole default config end to end (actors -> HBM staging with the cells riding in the
observation rows -> FIFO -> learner)."""
import copy

import numpy as np
import pytest
import torch

from helpers import ppo_configs, ref_state_dict

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
LOSS_STATS = ['_surr_loss', '_clip_surr_loss', '_kl_loss_adapt', '_entropy', '_pol_kl', '_val_loss', '_avg_return_targ',
              '_avg_log_sig', '_avg_behave_likelihood', '_avg_is_weight', '_ref_behave_diff', 'obs_running_mean',
              'obs_running_square', 'obs_running_std']


@pytest.mark.parametrize('B,L,D,H', [(8, 16, 17, 100), (5, 7, 11, 10), (64, 26, 17, 100)])
def test_lstm_forward_backward_match_torch(B, L, D, H):
    from surreal_b200.model.lstm_stem import LSTMStem, RnnTrainer
    g = torch.Generator().manual_seed(B + L + H)
    ref = torch.nn.LSTM(D, H, 1, batch_first=True)
    stem = LSTMStem(D, H, DEV)
    stem.load_torch({k: v.detach() for k, v in ref.state_dict().items()})
    x = torch.randn(B, L, D, generator=g)
    h0, c0 = torch.randn(1, B, H, generator=g) * 0.5, torch.randn(1, B, H, generator=g) * 0.5
    out, (hn, cn) = ref(x, (h0, c0))
    dout = torch.randn(B, L, H, generator=g)
    out.backward(dout)
    tr = RnnTrainer(stem, B, L)
    xf = torch.zeros(B * L, (D + 3) // 4 * 4, device=DEV)[:, :D]
    xf.copy_(x.reshape(B * L, D))
    h0d, c0d = h0[0].to(DEV).contiguous(), c0[0].to(DEV).contiguous()
    h = tr.forward(xf, h0d, c0d, H)
    torch.cuda.synchronize()
    assert float((h.cpu().reshape(B, L, H) - out.detach()).abs().max()) <= 1e-5
    dh = torch.zeros(B * L, (H + 3) // 4 * 4, device=DEV)
    dh[:, :H].copy_(dout.reshape(B * L, H))
    tr.backward(dh)
    torch.cuda.synchronize()
    gsum = tr.slabs.sum(0)
    wi, bi = stem.ih.get_layer(0, gsum[:stem.ih.size])
    wh, bh = stem.hh.get_layer(0, gsum[stem.ih.size:])
    scale = max(1.0, float(ref.weight_hh_l0.grad.abs().max()))
    assert float((wi.cpu() - ref.weight_ih_l0.grad).abs().max()) <= 2e-5 * scale
    assert float((wh.cpu() - ref.weight_hh_l0.grad).abs().max()) <= 2e-5 * scale
    assert float((bi.cpu() - ref.bias_ih_l0.grad).abs().max()) <= 2e-5 * scale
    assert float((bh.cpu() - ref.bias_hh_l0.grad).abs().max()) <= 2e-5 * scale


def _rnn_cfg(cfg, B=None):
    lc, ec, sc = ppo_configs(D=cfg['D'], A=cfg['A'], actor_h=cfg['actor_h'], critic_h=cfg['critic_h'], n_step=cfg['n_step'],
                             stride=cfg['n_step'], B=B or cfg['B'], mode=cfg['mode'], lr=cfg['lr'], exp_interval=cfg['exp_interval'])
    lc.algo.rnn.if_rnn_policy = True
    lc.algo.rnn.horizon, lc.algo.rnn.rnn_hidden, lc.algo.rnn.rnn_layer = cfg['horizon'], cfg['rnn_hidden'], cfg['rnn_layer']
    return lc, ec, sc


@pytest.mark.parametrize('tag', ['rnn_clip', 'rnn_adapt', 'rnn_adapt_biglr'])
def test_ppo_learn_rnn_matches_reference(golden, tag):
    from surreal_b200.learner import PPOLearner
    g = golden('ppo_learn_' + tag)
    cfg, hyper, stats = g.js('cfg'), g.js('hyper'), g.js('stats')
    lc, ec, sc = _rnn_cfg(cfg)
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ref_state_dict(g.sub('init/')))
    L.ref_target_model.update_target_params(L.model)
    lr, E = cfg['lr'], cfg['n_step'] - cfg['horizon'] + 1
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        st = L.learn({'obs': {'low_dim': {'flat_inputs': b['obs']}}, 'obs_next': {'low_dim': {'flat_inputs': b['obs_next']}},
                      'actions': b['actions'], 'rewards': b['rewards'], 'dones': b['dones'], 'persistent_infos': [b['pd']],
                      'onetime_infos': [b['h0'], b['c0']]})
        L.publish_parameter(it, message='')
        torch.cuda.synchronize()
        assert L._adv.shape == (cfg['B'], E)
        assert float((L._adv.cpu() - torch.tensor(b['adv'])).abs().max()) <= 1e-5
        rms = float(np.sqrt((b['ret'] ** 2).mean()))
        assert float((L._ret.cpu() - torch.tensor(b['ret'])).abs().max()) <= 1e-5 * max(1.0, rms)
        assert L.last_n_policy_epochs == hyper[it]['n_policy_epochs'], 'KL early stop diverged'
        for k, v in stats[it].items():
            assert k in st, k
            tol = 1e-5 * max(1.0, abs(v)) if k in LOSS_STATS else 2e-4 * max(1.0, abs(v))
            if k == '_val_explained_var':
                tol = 1e-4
            assert abs(st[k] - v) <= tol, '%s it%d: got %.9g expected %.9g' % (k, it, st[k], v)
        after = ref_state_dict(g.sub('it%d/after/' % it))
        got = L.model.state_dict()
        assert set(after) <= set(got), sorted(set(after) - set(got))
        worst = 0.0
        for k, e in after.items():
            gk = got[k].cpu().reshape(e.shape)
            if k.startswith('z_filter'):     # running sums of +-O(1) values: fp32 summation order, relative to the vector's scale
                assert float((gk - e).abs().max()) <= 1e-6 * max(1.0, float(e.abs().max())), k
                continue
            worst = max(worst, float((gk - e).abs().max()))
        assert worst <= max(2e-6, 0.02 * lr), 'params drifted by %.3e' % worst
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)
        assert L.exp_counter == hyper[it]['exp_counter']


def test_ppo_agent_act_rnn_matches_reference(golden):
    """One actor, seven steps: cells carried between steps, the PRE-step cells are the onetime_info, reset() zeroes them."""
    from surreal_b200.agent import PPOAgent
    g = golden('ppo_act_rnn')
    Hd = int(g['rnn_hidden'])
    cfg = dict(D=11, A=3, actor_h=[32, 24], critic_h=[28, 20], n_step=4, B=4, mode='clip', lr=1e-4, exp_interval=64, horizon=2,
               rnn_hidden=Hd, rnn_layer=1)
    lc, ec, sc = _rnn_cfg(cfg)
    ec.num_envs = 1
    ag = PPOAgent(lc, ec, sc, 0, 'training')
    ag.model.load_state_dict(ref_state_dict(g.sub('model/')))
    ag.set_noise(np.array([float(g['noise'])]))
    for i in range(len(g['obs'])):
        if i == int(g['reset_at']):
            ag.reset()
        a, info = ag.act(g['obs'][i], eps=g['eps'][i])
        np.testing.assert_allclose(info[0][0], g['h_before'][i], rtol=0, atol=2e-6)
        np.testing.assert_allclose(info[0][1], g['c_before'][i], rtol=0, atol=2e-6)
        np.testing.assert_allclose(info[1][0], g['pds'][i], rtol=0, atol=2e-6)
        np.testing.assert_allclose(a, g['actions'][i], rtol=0, atol=5e-6)


def test_default_ppo_config_runs_end_to_end():
    """BASELINE configs[0] stand-in: the UNMODIFIED PPO_DEFAULT_LEARNER_CONFIG (adapt mode + LSTM(100), n_step 25, stride 20,
    horizon 10, batch 64) on a synthetic 17-dim / 6-action env with the episode cap of 200: actors -> windows with the cells
    riding in the observation rows -> HBM FIFO -> learner.  The cells the learner reads for a window must be the cells the
    agent held at the window's first step, and learn() must agree with the RNN oracle on the same windows."""
    from surreal_b200.session import Config
    from surreal_b200.main.ppo_configs import (PPO_DEFAULT_LEARNER_CONFIG, PPO_DEFAULT_ENV_CONFIG, PPO_DEFAULT_SESSION_CONFIG,
                                               make_synthetic_env_config)
    from surreal_b200.agent import PPOAgent
    from surreal_b200.learner import PPOLearner
    from surreal_b200.replay import FIFOReplay
    from surreal_b200.env import SyntheticEnv
    from oracle.ppo_rnn import OraclePPOLearnerRNN
    from oracle.filters import ZFilter as OZ
    import tempfile
    lc = Config(copy.deepcopy(PPO_DEFAULT_LEARNER_CONFIG.to_dict()))
    ec = Config(copy.deepcopy(PPO_DEFAULT_ENV_CONFIG.to_dict()))
    sc = Config(copy.deepcopy(PPO_DEFAULT_SESSION_CONFIG.to_dict()))
    sc.folder = tempfile.mkdtemp()
    assert lc.algo.rnn.if_rnn_policy and lc.algo.ppo_mode == 'adapt'          # the reference's defaults, untouched
    N, D, A = 32, 17, 6            # 2 windows x 32 actors = one default batch of 64 (the default FIFO holds 96 + 3)
    make_synthetic_env_config(ec, N, D, A, seed=3)
    ec.limit_episode_length = 200
    n, stride, Hh, B = lc.algo.n_step, lc.algo.stride, lc.algo.rnn.rnn_hidden, lc.replay.batch_size
    R = FIFOReplay(lc, ec, sc)
    ag = PPOAgent(lc, ec, sc, 0, 'training')
    env = SyntheticEnv(N, D, A, limit_episode_length=200, seed=3)
    ag.env = w = ag.prepare_env_agent(env)
    obs, _ = w.reset()
    cells_at = {}
    for t in range(n + stride):                                    # two (overlapping) windows per actor
        cells_at[t] = (ag._h.clone(), ag._c.clone())
        a = ag.act(obs)
        assert torch.equal(a[1][0][0][:, 0], cells_at[t][0])      # onetime_info = cells BEFORE the step
        obs, _, _, _ = w.step(a)
    torch.cuda.synchronize()
    assert len(R) == 2 * N == B
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ag.model.state_dict())
    L.ref_target_model.update_target_params(L.model)
    b2 = R.sample(B)              # (step, actor) arrival order: rows 0..31 = first windows (cells: zeros), 32..63 = second ones
    h0, c0 = b2['onetime_infos'][0][:, 0], b2['onetime_infos'][1][:, 0]
    assert float(h0[:N].abs().max()) == 0.0 and float(c0[:N].abs().max()) == 0.0
    assert torch.equal(h0[N:], cells_at[stride][0]) and torch.equal(c0[N:], cells_at[stride][1])    # windows start at step `stride`
    assert float(h0[N:].abs().max()) > 0.0
    sd = {k: v.cpu() for k, v in L.model.state_dict().items()}
    al = [(sd['actor.model.seq.%d.weight' % (2 * i)], sd['actor.model.seq.%d.bias' % (2 * i)]) for i in range(3)]
    cl = [(sd['critic.model.seq.%d.weight' % (2 * i)], sd['critic.model.seq.%d.bias' % (2 * i)]) for i in range(3)]
    lstm = {k.split('.', 1)[1]: sd[k] for k in sd if k.startswith('rnn_stem.')}
    zf = OZ(D).load(sd['z_filter.running_sum'], sd['z_filter.running_sumsq'], sd['z_filter.count'])
    O = OraclePPOLearnerRNN(al, sd['actor.log_var'].view(1, A), cl, zf, lstm, A, n, B, lc.algo.rnn.horizon, Hh, 1, ppo_mode='adapt',
                            lr_actor=lc.algo.network.lr_actor, lr_critic=lc.algo.network.lr_critic)
    host = dict(obs=b2['obs']['low_dim']['flat_inputs'].cpu().numpy(), obs_next=b2['obs_next']['low_dim']['flat_inputs'].cpu().numpy(),
                actions=b2['actions'].cpu().numpy(), rewards=b2['rewards'].cpu().numpy(), dones=b2['dones'].cpu().numpy(),
                pd=b2['persistent_infos'][0].cpu().numpy(), h0=b2['onetime_infos'][0].cpu().numpy(), c0=b2['onetime_infos'][1].cpu().numpy())
    st = L.learn(b2)
    st_o = O.learn(host)
    torch.cuda.synchronize()
    assert L.last_n_policy_epochs == O.n_policy_epochs[-1]
    assert float((L._adv.cpu() - O.last_adv).abs().max()) <= 1e-5
    for k in ('_surr_loss', '_kl_loss_adapt', '_pol_kl', '_val_loss', '_entropy', '_avg_return_targ', '_avg_is_weight'):
        assert abs(st[k] - st_o[k]) <= 1e-5 * max(1.0, abs(st_o[k])), (k, st[k], st_o[k])
