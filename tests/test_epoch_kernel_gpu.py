"""The persistent learner kernel (csrc/epoch2.cu: all epochs of BOTH optimisers in ONE launch) against the launch chain it
replaces, on identical learners and batches: same statistics (1e-5), same number of policy epochs (incl. the KL early
stop), parameters and Adam moments equal to rounding-order noise, and bit-identical results run to run.  The reference
goldens (test_ppo_learner_gpu.py) and the full-size oracle comparisons (test_fullsize_gpu.py) run on the new path too,
since it is the default."""
import numpy as np
import pytest
import torch

from helpers import ppo_configs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GEN = '2'          # generation of the persistent learner kernel under test (csrc/epoch2.cu)


def _batch(B, n, D, A, seed):
    rng = np.random.default_rng(seed)
    obs = (rng.standard_normal((B, n, D)) * 1.2).astype(np.float32)
    obs_next = (rng.standard_normal((B, 1, D)) * 1.2).astype(np.float32)
    mean = np.tanh(rng.standard_normal((B, n, A)) * 0.3).astype(np.float32)
    std = np.exp(rng.uniform(-1.2, -0.6, (B, n, A))).astype(np.float32)
    pd = np.concatenate([mean, std], axis=2)
    actions = np.clip(rng.standard_normal((B, n, A)) * std + mean, -1, 1).astype(np.float32)
    rewards = (rng.standard_normal((B, n)) * 0.3).astype(np.float32)
    dones = np.zeros((B, n), dtype=np.float32)
    dones[rng.random(B) < 0.3, n - 1] = 1
    return {'obs': obs, 'obs_next': obs_next, 'actions': actions, 'rewards': rewards, 'dones': dones,
            'persistent_infos': [pd], 'onetime_infos': None}


def _pair(monkeypatch, **kw):
    from surreal_b200.learner import PPOLearner
    out = []
    for flag in ('0', GEN):
        monkeypatch.setenv('SB200_EPOCH_KERNEL', flag)
        torch.manual_seed(7)
        lc, ec, sc = ppo_configs(**kw)
        out.append(PPOLearner(lc, ec, sc))
    chain, fused = out
    assert not chain.use_epoch_kernel and fused.use_epoch_kernel and fused.epoch_kernel_gen == int(GEN)
    fused.model.actor.params.copy_(chain.model.actor.params)
    fused.model.critic.params.copy_(chain.model.critic.params)
    fused.ref_target_model.update_target_params(fused.model)
    chain.ref_target_model.update_target_params(chain.model)
    return chain, fused


CASES = [
    dict(D=11, A=3, actor_h=(32, 24), critic_h=(28, 20), n_step=6, stride=6, B=16, mode='clip', lr=1e-3),
    dict(D=11, A=3, actor_h=(32, 24), critic_h=(28, 20), n_step=6, stride=6, B=16, mode='adapt', lr=1e-3),
    dict(D=64, A=8, actor_h=(256, 256), critic_h=(256, 256), n_step=16, stride=16, B=1024, mode='clip', lr=1e-4),
    dict(D=64, A=8, actor_h=(256, 256), critic_h=(256, 256), n_step=16, stride=16, B=1024, mode='adapt', lr=1e-4),
    dict(D=17, A=6, actor_h=(300, 200), critic_h=(100, 68), n_step=8, stride=8, B=200, mode='clip', lr=3e-3, use_z=False),
    dict(D=20, A=1, actor_h=(64, 64), critic_h=(64, 64), n_step=5, stride=5, B=77, mode='adapt', lr=1e-2),
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_epoch_kernel_matches_launch_chain(case, monkeypatch):
    kw = CASES[case]
    chain, fused = _pair(monkeypatch, **kw)
    n_stop = 0
    for it in range(3):
        b = _batch(kw['B'], kw['n_step'], kw['D'], kw['A'], seed=100 * case + it)
        st_c = chain.learn(b)
        st_f = fused.learn(b)
        torch.cuda.synchronize()
        assert fused.last_n_policy_epochs == chain.last_n_policy_epochs, (it, fused.last_n_policy_epochs, chain.last_n_policy_epochs)
        n_stop += int(chain.last_n_policy_epochs < chain.epoch_policy)
        for k, v in st_c.items():
            assert abs(st_f[k] - v) <= (1e-5 if it == 0 else 1e-4) * max(1.0, abs(v)), (it, k, st_f[k], v)     # later calls start from parameters that differ by rounding order
        lr = kw['lr']
        for a, bb in ((chain.model.actor.params, fused.model.actor.params), (chain.model.critic.params, fused.model.critic.params)):
            d = (a - bb).abs()
            # Adam divides by sqrt(v): entries whose gradient is ~0 may flip sign between two summation orders, worth up to
            # one step each; everything else agrees to rounding
            assert float(d.median()) <= 1e-6 + 0.02 * lr, (it, float(d.median()))
            assert float(d.max()) <= 2.0 * lr * 10 * (it + 1), (it, float(d.max()))
        for a, bb in ((chain.actor_optim.exp_avg, fused.actor_optim.exp_avg), (chain.critic_optim.exp_avg, fused.critic_optim.exp_avg)):
            scale = float(a.abs().max()) + 1e-12
            assert float((a - bb).abs().max()) <= 1e-3 * scale + 1e-9, (it, float((a - bb).abs().max()), scale)
    assert fused._ek is not None and fused._ek[3] is not None, 'the one-launch kernel (epoch2.cu) was not the path taken'
    print('case %d: early stops %d of 3' % (case, n_stop))


def test_epoch_kernel_is_deterministic(monkeypatch):
    kw = CASES[2]
    from surreal_b200.learner import PPOLearner
    monkeypatch.setenv('SB200_EPOCH_KERNEL', GEN)
    res = []
    for rep in range(2):
        torch.manual_seed(3)
        lc, ec, sc = ppo_configs(**kw)
        L = PPOLearner(lc, ec, sc)
        for it in range(2):
            L.learn(_batch(kw['B'], kw['n_step'], kw['D'], kw['A'], seed=it))
        torch.cuda.synchronize()
        res.append((L.model.actor.params.clone(), L.model.critic.params.clone(), L._stats.clone()))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
