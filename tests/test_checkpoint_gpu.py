"""SURVEY §8(f) rank 3 on the REAL GPU learners: checkpoint -> restore -> continue (bit-identical with the optimiser
extension), and a checkpoint folder written by the REFERENCE's PPOLearner restored into surreal_b200.PPOLearner
(surreal/learner/ppo.py:668-678, surreal/learner/ddpg.py:383-387, surreal/utils/checkpoint.py:234-314)."""
import numpy as np
import pytest
import torch

from helpers import ppo_configs, ddpg_configs, ref_state_dict, ppo_batch

pytestmark = pytest.mark.gpu


def _ppo_batch(rng, B, n, D, A):
    pd = np.concatenate([np.tanh(rng.standard_normal((B, n, A))) * 0.5, np.full((B, n, A), 0.4)], -1).astype(np.float32)
    dones = np.zeros((B, n), dtype=np.float32)
    dones[rng.random(B) < 0.3, n - 1] = 1
    return {'obs': (rng.standard_normal((B, n, D)) * 1.3).astype(np.float32), 'obs_next': (rng.standard_normal((B, 1, D)) * 1.3).astype(np.float32),
            'actions': np.clip(rng.standard_normal((B, n, A)) * 0.4 + pd[:, :, :A], -1, 1), 'rewards': rng.standard_normal((B, n)) * 0.3,
            'dones': dones, 'persistent_infos': [pd], 'onetime_infos': None}


@pytest.mark.parametrize('mode', ['clip', 'adapt'])
def test_ppo_learner_checkpoint_roundtrip_continues_bit_identically(tmp_path, mode):
    from surreal_b200.learner import PPOLearner
    B, n, D, A = 64, 8, 12, 3

    def make(folder, restore_from=None):
        lc, ec, sc = ppo_configs(D=D, A=A, n_step=n, stride=n, B=B, mode=mode, lr=1e-3, exp_interval=B)
        sc.folder = str(folder)
        sc.checkpoint.learner.periodic = 1
        sc.checkpoint.learner.min_interval = 0
        sc.checkpoint.learner.include_optimizer = True
        if restore_from is not None:
            sc.checkpoint.restore = True
            sc.checkpoint.restore_folder = str(restore_from)
        return PPOLearner(lc, ec, sc)
    rng = np.random.default_rng(3)
    torch.manual_seed(3)
    A_ = make(tmp_path / 'a')
    for it in range(3):
        A_.learn(_ppo_batch(rng, B, n, D, A))                   # learn() writes the checkpoint (ppo.py:607) ...
        if it < 2:
            A_.publish_parameter(it)                           # ... BEFORE the publish-time adaptation (clip_epsilon / beta,
                                                               # ref-model refresh, LR schedule) of the same iteration
    torch.manual_seed(999)                                      # the fresh learner starts from DIFFERENT weights
    B_ = make(tmp_path / 'b', restore_from=tmp_path / 'a')
    assert B_.current_iteration == A_.current_iteration == 3
    for k, v in A_.model.state_dict().items():
        assert torch.equal(v, B_.model.state_dict()[k]), k
    for k, v in A_.ref_target_model.state_dict().items():
        assert torch.equal(v, B_.ref_target_model.state_dict()[k]), k
    assert B_.actor_lr_scheduler.n_step == A_.actor_lr_scheduler.n_step and B_.actor_lr_scheduler.get_lr() == A_.actor_lr_scheduler.get_lr()
    assert getattr(B_, 'clip_epsilon', None) == getattr(A_, 'clip_epsilon', None) and getattr(B_, 'beta', None) == getattr(A_, 'beta', None)
    nxt = _ppo_batch(rng, B, n, D, A)
    sa, sb = A_.learn(nxt), B_.learn(nxt)
    torch.cuda.synchronize()
    for k in sa:
        assert sa[k] == sb[k], (k, sa[k], sb[k])
    assert torch.equal(A_.model.actor.params, B_.model.actor.params) and torch.equal(A_.model.critic.params, B_.model.critic.params)
    assert torch.equal(A_.model.z_stats, B_.model.z_stats)
    assert torch.equal(A_.actor_optim.exp_avg_sq, B_.actor_optim.exp_avg_sq)


def test_ddpg_learner_checkpoint_roundtrip_continues_bit_identically(tmp_path):
    from surreal_b200.learner import DDPGLearner
    B, D, A = 32, 9, 3

    def make(folder, restore_from=None):
        lc, ec, sc = ddpg_configs(D=D, A=A, B=B, target={'type': 'soft', 'tau': 0.05, 'interval': 1})
        sc.folder = str(folder)
        sc.checkpoint.learner.periodic = 1
        sc.checkpoint.learner.min_interval = 0
        sc.checkpoint.learner.include_optimizer = True
        if restore_from is not None:
            sc.checkpoint.restore = True
            sc.checkpoint.restore_folder = str(restore_from)
        return DDPGLearner(lc, ec, sc)
    rng = np.random.default_rng(4)
    mk = lambda: {'obs': {'low_dim': {'flat_inputs': rng.standard_normal((B, D)).astype(np.float32)}},   # noqa: E731
                  'obs_next': {'low_dim': {'flat_inputs': rng.standard_normal((B, D)).astype(np.float32)}},
                  'actions': rng.uniform(-1, 1, (B, A)).astype(np.float32), 'rewards': rng.standard_normal((B, 1)),
                  'dones': (rng.random((B, 1)) < 0.1).astype(np.float64)}
    torch.manual_seed(4)
    A_ = make(tmp_path / 'a')
    for _ in range(3):
        A_.learn(mk())
    torch.manual_seed(777)
    B_ = make(tmp_path / 'b', restore_from=tmp_path / 'a')
    assert B_.current_iteration == 3
    nxt = mk()
    sa, sb = A_.learn(nxt), B_.learn(nxt)
    torch.cuda.synchronize()
    for k in ('actor_loss', 'critic_loss', 'Q_target', 'Q_policy'):
        assert sa[k] == sb[k], (k, sa[k], sb[k])
    for x, y in ((A_.model, B_.model), (A_.model_target, B_.model_target)):
        assert torch.equal(x.actor.params, y.actor.params) and torch.equal(x.critic.params, y.critic.params)


def test_reference_written_ppo_learner_checkpoint_restores_into_gpu_learner(tmp_path, golden):
    """The folder in tests/golden/ppo_learner_ckpt.npz was written by the reference's PeriodicCheckpoint tracking a real
    reference PPOLearner.  Restored into surreal_b200.PPOLearner: same model / ref model / iteration, the pickled
    scheduler objects map onto our schedule, and the next learn() reproduces what the REFERENCE computed after restoring
    the same folder (fresh Adam state on both sides: the reference does not checkpoint its optimisers)."""
    from surreal_b200.learner import PPOLearner
    g = golden('ppo_learner_ckpt')
    cfg = g.js('cfg')
    ck = tmp_path / 'ref' / 'checkpoint'
    ck.mkdir(parents=True)
    names = g.js('file_names') if g['file_names'].ndim == 0 else [str(x) for x in g['file_names']]
    for fn in names:
        (ck / fn).write_bytes(bytes(g['file/' + fn]))
    lc, ec, sc = ppo_configs(D=cfg['D'], A=cfg['A'], actor_h=cfg['actor_h'], critic_h=cfg['critic_h'], n_step=cfg['n_step'],
                             stride=cfg['n_step'], B=cfg['B'], mode='clip', lr=cfg['lr'], exp_interval=cfg['exp_interval'])
    sc.folder = str(tmp_path / 'mine')
    sc.checkpoint.restore = True
    sc.checkpoint.restore_folder = str(tmp_path / 'ref')
    L = PPOLearner(lc, ec, sc)
    assert L.current_iteration == cfg['current_iteration']
    assert L.actor_lr_scheduler.n_step == cfg['sched_n_step']
    for name, model in (('saved/model/', L.model), ('saved/ref/', L.ref_target_model)):
        exp = ref_state_dict(g.sub(name))
        got = model.state_dict()
        for k, e in exp.items():
            assert torch.equal(got[k].cpu().reshape(e.shape), e), k
    st = L.learn(ppo_batch(g.sub('next/')))
    torch.cuda.synchronize()
    ref = g.js('next_stats')
    for k in ('_surr_loss', '_clip_surr_loss', '_entropy', '_pol_kl', '_val_loss', '_avg_return_targ', '_avg_is_weight',
              '_ref_behave_diff', 'grad_norm_actor', 'grad_norm_critic'):
        assert abs(st[k] - ref[k]) <= 1e-5 * max(1.0, abs(ref[k])), (k, st[k], ref[k])
    exp = ref_state_dict(g.sub('next/after/'))
    got = L.model.state_dict()
    for k, e in exp.items():
        d = float((got[k].cpu().reshape(e.shape) - e).abs().max())
        if k.startswith('z_filter'):
            assert d <= 2e-7 * float(e.abs().max()), k          # running sums ~2e2: one fp32 ulp of summation order
        else:
            assert d <= 2e-6, k                                   # lr 1e-4: 2 % of one Adam step
