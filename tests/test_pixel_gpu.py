"""Pixel path (BASELINE configs[3]): uint8 frames -> CNN stem shared by actor and critic and
trained by BOTH optimisers (ppo_net.py:136-140,202-224,268-273; builders.py:8-33).

  * PPOLearner.learn + publish against goldens produced by the REFERENCE's PPOLearner in pixel mode;
  * actors -> HBM staging -> FIFO -> learner on the device pixel env: frames stay uint8 and bit-identical end to end, the
    behaviour policy rows equal the oracle's stem + actor on the staged frames."""
import numpy as np
import pytest
import torch

from helpers import ppo_configs, ref_state_dict

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

LOSS_STATS = ['_surr_loss', '_clip_surr_loss', '_kl_loss_adapt', '_entropy', '_pol_kl', '_val_loss', '_avg_return_targ',
              '_avg_log_sig', '_avg_behave_likelihood', '_avg_is_weight', '_ref_behave_diff']


def _pixel_cfg(C, HW, F, A, actor_h, critic_h, n, B, mode, lr, exp_interval, N=None):
    lc, ec, sc = ppo_configs(D=1, A=A, actor_h=actor_h, critic_h=critic_h, n_step=n, stride=n, B=B, mode=mode, lr=lr,
                             exp_interval=exp_interval, use_z=False)
    ec.obs_spec = {'pixel': {'camera0': (C, HW, HW)}}
    ec.pixel_input = True
    lc.model.cnn_feature_dim = F
    if N is not None:
        ec.num_envs = N
    return lc, ec, sc


@pytest.mark.parametrize('tag', ['pixel_clip', 'pixel_adapt'])
def test_ppo_learn_pixel_matches_reference(golden, tag):
    from surreal_b200.learner import PPOLearner
    g = golden('ppo_learn_' + tag)
    cfg, hyper, stats = g.js('cfg'), g.js('hyper'), g.js('stats')
    lc, ec, sc = _pixel_cfg(cfg['C'], cfg['HW'], cfg['cnn_feature_dim'], cfg['A'], cfg['actor_h'], cfg['critic_h'], cfg['n_step'],
                            cfg['B'], cfg['mode'], cfg['lr'], cfg['exp_interval'])
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ref_state_dict(g.sub('init/')))
    L.ref_target_model.update_target_params(L.model)
    lr = cfg['lr']
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        assert b['obs'].dtype == np.uint8
        st = L.learn({'obs': {'pixel': {'camera0': b['obs']}}, 'obs_next': {'pixel': {'camera0': b['obs_next']}},
                      'actions': b['actions'], 'rewards': b['rewards'], 'dones': b['dones'], 'persistent_infos': [b['pd']],
                      'onetime_infos': None})
        L.publish_parameter(it, message='')
        torch.cuda.synchronize()
        assert float((L._adv.cpu() - torch.tensor(b['adv'])).abs().max()) <= 1e-5
        rms = float(np.sqrt((b['ret'] ** 2).mean()))
        assert float((L._ret.cpu() - torch.tensor(b['ret'])).abs().max()) <= 1e-5 * max(1.0, rms)
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
            worst = max(worst, float((got[k].cpu().reshape(e.shape) - e).abs().max()))
        assert worst <= max(2e-6, 0.02 * lr), 'params drifted by %.3e' % worst
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)


def test_pixel_actors_replay_learner_end_to_end():
    """Device pixel env -> PPOAgent.act (stem + head + sampling) -> window staging -> HBM FIFO -> PPOLearner.learn.
    Frames must arrive in the learner's batch bit-identical to what the env produced; pd rows must equal the oracle's
    stem + actor forward on those frames; learn() must match the pixel oracle learner on the same windows."""
    import torch.nn.functional as F
    from surreal_b200.agent import PPOAgent
    from surreal_b200.learner import PPOLearner
    from surreal_b200.replay import FIFOReplay
    from surreal_b200.env import SyntheticPixelEnv
    from oracle.ppo_pixel import OraclePPOLearnerPixel
    N, C, HW, Fd, A, n = 16, 4, 28, 24, 3, 5
    lc, ec, sc = _pixel_cfg(C, HW, Fd, A, (32, 24), (28, 20), n, N, 'clip', 1e-4, N, N=N)
    lc.replay.memory_size = 4 * N
    R = FIFOReplay(lc, ec, sc)
    ag = PPOAgent(lc, ec, sc, 0, 'training')
    ag.set_noise(np.linspace(-0.2, 0.2, N))
    env = SyntheticPixelEnv(N, (C, HW, HW), A, limit_episode_length=2 * n, seed=5)
    ag.env = w = ag.prepare_env_agent(env)
    obs, _ = w.reset()
    seen = [obs['pixel']['camera0'].clone()]
    for _ in range(n):
        a = ag.act(obs)
        obs, _, _, info = w.step(a)
        seen.append(info['obs_next']['pixel']['camera0'].clone())
    torch.cuda.synchronize()
    assert len(R) == N
    batch = R.sample(N)
    fr = torch.cat([batch['obs']['pixel']['camera0'], batch['obs_next']['pixel']['camera0']], 1)     # [N, n+1, C, H, W] uint8
    assert fr.dtype == torch.uint8
    for t in range(n + 1):
        assert torch.equal(fr[:, t], seen[t]), 'frame %d changed on the way through staging / replay' % t
    # oracle stem + actor on the staged frames == behaviour policy rows
    sd = {k: v.cpu() for k, v in ag.model.state_dict().items()}
    conv = [(sd['cnn_stem.model.seq.0.weight'], sd['cnn_stem.model.seq.0.bias']), (sd['cnn_stem.model.seq.2.weight'], sd['cnn_stem.model.seq.2.bias'])]
    fc = (sd['cnn_stem.model.seq.5.weight'], sd['cnn_stem.model.seq.5.bias'])
    al = [(sd['actor.model.seq.%d.weight' % (2 * i)], sd['actor.model.seq.%d.bias' % (2 * i)]) for i in range(3)]
    cl = [(sd['critic.model.seq.%d.weight' % (2 * i)], sd['critic.model.seq.%d.bias' % (2 * i)]) for i in range(3)]
    with torch.no_grad():
        h = fr[:, :n].reshape(-1, C, HW, HW).float().cpu() / 255.0
        h = torch.relu(F.conv2d(h, conv[0][0], conv[0][1], stride=4))
        h = torch.relu(F.conv2d(h, conv[1][0], conv[1][1], stride=2)).flatten(1)
        h = torch.relu(F.linear(h, fc[0], fc[1]))
        for i, (wt, bt) in enumerate(al):
            h = F.linear(h, wt, bt)
            h = torch.relu(h) if i < 2 else torch.tanh(h)
    pd = batch['persistent_infos'][0].cpu().view(N * n, 2 * A)
    assert float((pd[:, :A] - h).abs().max()) <= 1e-5
    std = torch.exp(sd['actor.log_var']).view(1, A) * torch.exp(torch.tensor(np.linspace(-0.2, 0.2, N), dtype=torch.float32)).repeat_interleave(n).view(-1, 1)
    assert float((pd[:, A:] - std).abs().max()) <= 1e-6
    # learn() on the HBM batch vs the pixel oracle on the same windows
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ag.model.state_dict())
    L.ref_target_model.update_target_params(L.model)
    O = OraclePPOLearnerPixel(al, sd['actor.log_var'].view(1, A), cl, conv, [4, 2], fc, A, n, N, ppo_mode='clip', lr_actor=1e-4,
                              lr_critic=1e-4, exp_interval=N)
    host = dict(obs=fr[:, :n].cpu().numpy(), obs_next=fr[:, n:].cpu().numpy(), actions=batch['actions'].cpu().numpy(),
                rewards=batch['rewards'].cpu().numpy(), dones=batch['dones'].cpu().numpy(), pd=batch['persistent_infos'][0].cpu().numpy())
    st = L.learn(batch)
    st_o = O.learn(host)
    torch.cuda.synchronize()
    assert L.last_n_policy_epochs == O.n_policy_epochs[-1]
    for k in ('_surr_loss', '_clip_surr_loss', '_pol_kl', '_val_loss', '_entropy', '_avg_return_targ', '_avg_is_weight'):
        assert abs(st[k] - st_o[k]) <= 1e-5 * max(1.0, abs(st_o[k])), (k, st[k], st_o[k])
    assert float((L._adv.cpu().view(-1) - O.last_adv.view(-1)).abs().max()) <= 1e-5
