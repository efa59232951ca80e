"""PPOLearner (CUDA) vs golden vectors produced by the reference's own PPOLearner.learn / publish."""
import numpy as np
import pytest
import torch

from helpers import ppo_configs, ref_state_dict, ppo_batch

pytestmark = pytest.mark.gpu

# reference stat name -> tolerance class
LOSS_STATS = ['_surr_loss', '_clip_surr_loss', '_kl_loss_adapt', '_entropy', '_pol_kl', '_val_loss',
              '_avg_return_targ', '_avg_log_sig', '_avg_behave_likelihood', '_avg_is_weight', '_ref_behave_diff',
              'obs_running_mean', 'obs_running_square', 'obs_running_std', 'reward_mean']


def _make(cfg, g):
    from surreal_b200.learner import PPOLearner
    lc, ec, sc = ppo_configs(D=cfg['D'], A=cfg['A'], actor_h=cfg['actor_h'], critic_h=cfg['critic_h'],
                             n_step=cfg['n_step'], stride=cfg['n_step'], B=cfg['B'], mode=cfg['mode'], lr=cfg['lr'],
                             exp_interval=cfg['exp_interval'], use_r=cfg['use_r_filter'],
                             reward_scale=cfg['reward_scale'])
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ref_state_dict(g.sub('init/')))
    L.ref_target_model.update_target_params(L.model)
    return L


@pytest.mark.parametrize('tag', ['clip', 'adapt', 'clip_biglr', 'adapt_biglr', 'clip_rfilter'])
def test_ppo_learn_matches_reference(golden, tag):
    g = golden('ppo_learn_' + tag)
    cfg, hyper, stats = g.js('cfg'), g.js('hyper'), g.js('stats')
    L = _make(cfg, g)
    lr = cfg['lr']
    for it in range(cfg['iters']):
        b = g.sub('it%d/' % it)
        st = L.learn(ppo_batch(b))
        L.publish_parameter(it, message='')
        torch.cuda.synchronize()
        assert L.last_n_policy_epochs == hyper[it]['n_policy_epochs'], 'KL early stop diverged'
        for k, v in stats[it].items():
            assert k in st, k
            tol = 1e-5 * max(1.0, abs(v)) if k in LOSS_STATS else 2e-4 * max(1.0, abs(v))
            if k == '_val_explained_var':
                tol = 1e-4
            assert abs(st[k] - v) <= tol, '%s it%d: got %.9g expected %.9g' % (k, it, st[k], v)
        after = ref_state_dict(g.sub('it%d/after/' % it))
        got = L.model.state_dict()
        worst = 0.0
        for k, e in after.items():
            gk = got[k].cpu().reshape(e.shape)
            if k.startswith('z_filter'):      # running sums (~1e2): fp32 accumulation order -> a few ulp, relative bar
                assert float(((gk - e).abs() / e.abs().clamp_min(1.0)).max()) <= 1e-6, k
                continue
            worst = max(worst, float((gk - e).abs().max()))
        # 20 Adam steps of size <= lr each: agree to a small fraction of ONE step
        assert worst <= max(2e-6, 0.02 * lr), 'params drifted by %.3e' % worst
        if cfg['mode'] == 'clip':
            assert L.clip_epsilon == pytest.approx(hyper[it]['clip_epsilon'], rel=1e-12)
        else:
            assert L.beta == pytest.approx(hyper[it]['beta'], rel=1e-12)
        assert L.exp_counter == hyper[it]['exp_counter']
        ref_after = ref_state_dict(g.sub('it%d/ref/' % it))
        got_ref = L.ref_target_model.state_dict()
        for k in ['actor.log_var', 'z_filter.count', 'actor.model.seq.0.weight']:
            assert float((got_ref[k].cpu().reshape(ref_after[k].shape) - ref_after[k]).abs().max()) <= max(2e-6, 0.02 * lr)


def test_gae_and_values_inside_learner(golden):
    """The learner's fused critic pass + GAE against the reference's _gae_and_return (values through the real critic)."""
    from surreal_b200.learner import PPOLearner
    g = golden('gae_mlp')
    B, n, D = g['obs'].shape
    lc, ec, sc = ppo_configs(D=D, A=3, n_step=n, stride=n, B=B)
    L = PPOLearner(lc, ec, sc)
    L.model.load_state_dict(ref_state_dict(g.sub('model/')))
    batch = {'obs': g['obs'], 'obs_next': g['obs_next'], 'actions': np.zeros((B, n, 3)), 'rewards': g['rewards'],
             'dones': g['dones'], 'persistent_infos': [np.ones((B, n, 6), dtype=np.float32)]}
    L._preprocess_batch_ppo(batch)
    adv, ret = L._gae_and_return()
    torch.cuda.synchronize()
    v = L._values.view(B, n + 1).cpu()
    assert float((v - torch.tensor(g['values_raw'])).abs().max()) <= 1e-5 * max(1.0, float(np.sqrt((g['values_raw'] ** 2).mean())))
    assert float((adv.cpu() - torch.tensor(g['adv'])).abs().max()) <= 1e-5
    assert float((ret.cpu() - torch.tensor(g['ret'])).abs().max()) <= 1e-5 * max(1.0, float(np.sqrt((g['ret'] ** 2).mean())))
