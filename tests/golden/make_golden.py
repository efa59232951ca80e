#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference (SurrealAI/surreal @ /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference's third-party plumbing is replaced by the stand-ins of ``_ref_harness.py``;
every number stored here is produced by the reference's own functions:

  fixture               reference code that produced it
  --------------------  -----------------------------------------------------------------
  pd.npz                surreal/model/ppo_net.py:29-91   DiagGauss.{loglikelihood,likelihood,kl,entropy}
  zfilter.npz           surreal/model/z_filter.py:44-79  ZFilter.{forward,z_update}
  rfilter.npz           surreal/model/reward_filter.py:34-57
  gae_*.npz             surreal/learner/ppo.py:355-418   PPOLearner._gae_and_return (MLP + RNN mode)
  ppo_learn_*.npz       surreal/learner/ppo.py:420-666   PPOLearner.learn / publish_parameter / _post_publish
  ppo_learn_rnn_*.npz   the same in RNN mode (LSTM stem, horizon GAE: ppo.py:389-406,507-525, ppo_net.py:143-152)
  ppo_learn_pixel_*.npz the same in pixel mode (shared CNN stem: ppo_net.py:136-140,268-273, builders.py:8-33)
  ddpg_optimize_*.npz   surreal/learner/ddpg.py:186-428  DDPGLearner.preprocess/_optimize/_target_update
  ddpg_optimize_td3_*   the same with use_double_critic / use_action_regularization (ddpg.py:267-283,298-321)
  replay.npz            surreal/replay/{fifo,uniform}_replay.py  insert / sample / start_sample_condition
  window_*.npz          surreal/env/exp_sender_wrapper.py:72-112,153-264
  aggregate.npz         surreal/learner/aggregator.py:33-103,106-262
  ppo_act.npz           surreal/agent/ppo_agent.py:106-154
  ppo_act_rnn.npz       surreal/agent/ppo_agent.py:84-93,133-137,169-183  (LSTM cells hand-off, reset)
  ddpg_act.npz          surreal/agent/ddpg_agent.py:155-184 + action_noise.py:9-39
  ddpg_act_ou.npz       the same with OrnsteinUhlenbeckActionNoise (action_noise.py:22-39) and pre_episode reset
  checkpoint.npz        surreal/utils/checkpoint.py:18-347  files written by PeriodicCheckpoint
  configs.npz           the default config trees of main/{ppo,ddpg}_configs.py + session/default_configs.py

All inputs are seeded here and stored next to the outputs, so the fixtures are self-contained.
"""
import copy
import json
import os
import random
import sys
import tempfile
import warnings

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_harness as H  # noqa: E402

H.install()

import numpy as np  # noqa: E402
import torch  # noqa: E402

from surreal.session import Config  # noqa: E402
from surreal.main.ppo_configs import (PPO_DEFAULT_LEARNER_CONFIG, PPO_DEFAULT_ENV_CONFIG,  # noqa: E402
                                      PPO_DEFAULT_SESSION_CONFIG)
from surreal.main.ddpg_configs import (DDPG_DEFAULT_LEARNER_CONFIG, DDPG_DEFAULT_ENV_CONFIG,  # noqa: E402
                                       DDPG_DEFAULT_SESSION_CONFIG)
from surreal.learner.ppo import PPOLearner  # noqa: E402
from surreal.learner.ddpg import DDPGLearner  # noqa: E402
from surreal.learner.aggregator import MultistepAggregatorWithInfo, SSARAggregator  # noqa: E402
from surreal.model.ppo_net import DiagGauss  # noqa: E402
from surreal.model.z_filter import ZFilter  # noqa: E402
from surreal.model.reward_filter import RewardFilter  # noqa: E402
from surreal.replay import FIFOReplay, UniformReplay  # noqa: E402
from surreal.agent.ppo_agent import PPOAgent  # noqa: E402
from surreal.agent.ddpg_agent import DDPGAgent  # noqa: E402
import surreal.env.exp_sender_wrapper as ESW  # noqa: E402

torch.set_num_threads(1)

# Python >= 3.11 turns `continuous = ()` / `discrete = ()` of surreal/env/base.py:7-9 into ALIASES of
# one member (the StringEnum trick of utils/common.py:85-86 assigns _value_ too late), which would
# wrongly send continuous actions down aggregator.py:172's broken discrete branch.  Interpreter rot,
# not reference behaviour: restore two distinct members for the aggregators.
import enum  # noqa: E402
import surreal.learner.aggregator as _AG  # noqa: E402


class _ActionType(enum.Enum):
    continuous = 'continuous'
    discrete = 'discrete'


_AG.ActionType = _ActionType


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        if isinstance(v, (dict, list, str)) and not isinstance(v, np.ndarray):
            v = np.array(json.dumps(v))
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('wrote %-28s %7.1f KB' % (name + '.npz', os.path.getsize(path) / 1024))


def sd_np(module, prefix=''):
    return {prefix + k.replace('.', '/'): v.detach().cpu().numpy().copy()
            for k, v in module.state_dict().items()}


# --------------------------------------------------------------------------------------
def cfg_ppo(D=11, A=3, actor_h=(32, 24), critic_h=(28, 20), n_step=6, stride=6, B=16,
            mode='clip', rnn=False, horizon=3, rnn_hidden=10, lr=1e-4, exp_interval=4096,
            use_z=True, use_r=False, reward_scale=1.0, norm_adv=True):
    lc = Config(copy.deepcopy(PPO_DEFAULT_LEARNER_CONFIG.to_dict()))
    ec = Config(copy.deepcopy(PPO_DEFAULT_ENV_CONFIG.to_dict()))
    sc = Config(copy.deepcopy(PPO_DEFAULT_SESSION_CONFIG.to_dict()))
    sc.folder = tempfile.mkdtemp()
    ec.obs_spec = {'low_dim': {'flat_inputs': (D,)}}
    ec.action_spec = {'dim': (A,), 'type': 'continuous'}
    lc.model.actor_fc_hidden_sizes = list(actor_h)
    lc.model.critic_fc_hidden_sizes = list(critic_h)
    lc.algo.ppo_mode = mode
    lc.algo.rnn.if_rnn_policy = rnn
    lc.algo.rnn.horizon = horizon
    lc.algo.rnn.rnn_hidden = rnn_hidden
    lc.algo.n_step = n_step
    lc.algo.stride = stride
    lc.algo.network.lr_actor = lr
    lc.algo.network.lr_critic = lr
    lc.algo.use_z_filter = use_z
    lc.algo.use_r_filter = use_r
    lc.algo.advantage.reward_scale = reward_scale
    lc.algo.advantage.norm_adv = norm_adv
    lc.replay.batch_size = B
    lc.parameter_publish.exp_interval = exp_interval
    return lc, ec, sc


def make_ppo_batch(L, rng, B, n, D, A, log_sig_range=0.25, done_last_frac=0.25):
    """Synthetic windows in the aggregator's output format (aggregator.py:151-183)."""
    obs = rng.standard_normal((B, n, D)).astype(np.float32) * 1.5 + 0.3
    obs_next = rng.standard_normal((B, 1, D)).astype(np.float32) * 1.5 + 0.3
    with torch.no_grad():
        flat = {'low_dim': {'flat_inputs': torch.tensor(obs.reshape(B * n, D))}}
        if L.if_rnn_policy:
            flat = {'low_dim': {'flat_inputs': torch.tensor(obs)}}
            h0 = torch.zeros(1, B, L.learner_config.algo.rnn.rnn_hidden)
            pd = L.model.forward_actor(flat, (h0, h0.clone())).numpy().reshape(B, n, 2 * A)
        else:
            pd = L.model.forward_actor(flat).numpy().reshape(B, n, 2 * A)
    noise = rng.uniform(-log_sig_range, log_sig_range, size=(B, 1, 1))
    pd = pd.copy()
    pd[:, :, A:] *= np.exp(noise).astype(np.float32)           # ppo_agent.py:139
    eps = rng.standard_normal((B, n, A))                        # float64, as np.random.randn
    actions = np.clip(eps * pd[:, :, A:] + pd[:, :, :A], -1, 1)  # float64 (ppo_net.py:83)
    rewards = rng.standard_normal((B, n)) * 0.5 + 0.1           # python floats -> float64
    dones = np.zeros((B, n), dtype=np.float32)
    dones[rng.random(B) < done_last_frac, n - 1] = 1.0          # only the last step can be done
    batch = H._AttrDict(
        obs={'low_dim': {'flat_inputs': obs.copy()}},
        obs_next={'low_dim': {'flat_inputs': obs_next.copy()}},
        actions=actions.copy(), rewards=rewards.copy(), dones=dones.copy(),
        persistent_infos=[pd.astype(np.float32).copy()], onetime_infos=None)
    if L.if_rnn_policy:
        Hd = L.learner_config.algo.rnn.rnn_hidden
        h = (rng.standard_normal((B, 1, Hd)) * 0.1).astype(np.float32)
        c = (rng.standard_normal((B, 1, Hd)) * 0.1).astype(np.float32)
        batch.onetime_infos = [h.copy(), c.copy()]
    raw = dict(obs=obs, obs_next=obs_next, actions=actions, rewards=rewards, dones=dones,
               pd=pd.astype(np.float32))
    if L.if_rnn_policy:
        raw.update(h0=h, c0=c)
    return batch, raw


def gen_pd():
    rng = np.random.default_rng(100)
    A = 5
    pdc = DiagGauss(A)
    a = torch.tensor(rng.uniform(-1, 1, (33, A)).astype(np.float32))
    p0 = torch.tensor(np.concatenate([rng.uniform(-1, 1, (33, A)), rng.uniform(0.2, 1.5, (33, A))], 1).astype(np.float32))
    p1 = torch.tensor(np.concatenate([rng.uniform(-1, 1, (33, A)), rng.uniform(0.2, 1.5, (33, A))], 1).astype(np.float32))
    # a few rows far in the tail so that the 1e-5 likelihood clamp engages
    a[:3] = a[:3] * 0 + 1.0
    p0[:3, :A] = -1.0
    p0[:3, A:] = 0.05
    save('pd', a=a, p0=p0, p1=p1,
         loglik=pdc.loglikelihood(a, p0), lik=pdc.likelihood(a, p0),
         kl01=pdc.kl(p0, p1), ent=pdc.entropy(p0))


def gen_filters():
    rng = np.random.default_rng(101)
    D = 7
    zf = ZFilter({'low_dim': {'a': (3,), 'b': (4,)}})
    x0 = torch.tensor(rng.standard_normal((5, D)).astype(np.float32))
    y_init = zf.forward(x0)                    # with the eps-initialised statistics
    xs = [torch.tensor((rng.standard_normal((19, D)) * 3 + 1).astype(np.float32)) for _ in range(3)]
    outs, states = [], []
    for x in xs:
        zf.z_update(x)
        states.append(np.concatenate([zf.running_sum.numpy(), zf.running_sumsq.numpy(), zf.count.numpy()]))
        outs.append(zf.forward(x0).numpy().copy())
    save('zfilter', x0=x0, y_init=y_init, xs=torch.stack(xs), outs=np.stack(outs), states=np.stack(states),
         running_mean=zf.running_mean(), running_std=zf.running_std(), running_square=zf.running_square())

    rf = RewardFilter()
    r = [torch.tensor((rng.standard_normal((4, 6)) * 2 + 0.5).astype(np.float32)) for _ in range(3)]
    outs, states = [], []
    for x in r:
        outs.append(rf.forward(x).numpy().copy())          # ppo.py:454-455: forward THEN update
        rf.update(x)
        states.append([rf.count.item(), rf.running_sum.item(), rf.running_sumsq.item()])
    save('rfilter', r=torch.stack(r), outs=np.stack(outs), states=np.array(states, dtype=np.float64),
         reward_mean=rf.reward_mean())


def gen_gae():
    # --- MLP mode, through the real critic
    for tag, kw, general_dones in [('mlp', dict(n_step=9, B=12, norm_adv=True), False),
                                   ('mlp_nonorm', dict(n_step=5, B=4, norm_adv=False), True),
                                   ('rnn', dict(n_step=9, B=6, rnn=True, horizon=4, rnn_hidden=10), False)]:
        torch.manual_seed(7)
        lc, ec, sc = cfg_ppo(**kw)
        L = H.construct_without_initialize(PPOLearner, lc, ec, sc)
        rng = np.random.default_rng(7)
        B, n, D = L.batch_size, L.n_step, 11
        # give the z-filter non-trivial statistics
        L.model.z_filter.z_update(torch.tensor((rng.standard_normal((50, D)) * 2 + 0.5).astype(np.float32)))
        batch, raw = make_ppo_batch(L, rng, B, n, D, 3)
        if general_dones:
            raw['dones'][1, 2] = 1.0           # mask in the middle of a window (general elementwise mask)
            batch.dones = raw['dones'].copy()
        pre = L._preprocess_batch_ppo(batch)
        captured = {}
        orig = L.model.forward_critic

        def hook(obs, cells=None):
            v = orig(obs, cells)
            captured['values_raw'] = v.detach().clone()
            return v
        L.model.forward_critic = hook
        if L.if_rnn_policy:
            h = pre.onetime_infos[0].transpose(0, 1).contiguous()
            c = pre.onetime_infos[1].transpose(0, 1).contiguous()
            L.cells = (h, c)
        adv, ret = L._gae_and_return(pre.obs, pre.obs_next, pre.rewards, pre.dones)
        extra = {}
        if L.if_rnn_policy:
            extra = dict(h0=raw['h0'], c0=raw['c0'], horizon=L.horizon)
        save('gae_' + tag, obs=raw['obs'], obs_next=raw['obs_next'], rewards=raw['rewards'],
             dones=raw['dones'], values_raw=captured['values_raw'].reshape(B, n + 1),
             adv=adv, ret=ret, gamma=L.gamma, lam=L.lam, n_step=n, norm_adv=int(L.norm_adv),
             **sd_np(L.model, 'model/'), **extra)


def gen_ppo_learn():
    variants = [
        ('clip', dict(mode='clip', B=16, n_step=6, exp_interval=32), 3),
        ('adapt', dict(mode='adapt', B=16, n_step=6, exp_interval=32), 3),
        # large LR: KL early stop (ppo.py:556) and the adapt-mode cutoff branch (ppo.py:275) both fire
        ('clip_biglr', dict(mode='clip', B=16, n_step=6, lr=3e-3, exp_interval=16), 3),
        ('adapt_biglr', dict(mode='adapt', B=16, n_step=6, lr=3e-3, exp_interval=16), 3),
        ('clip_rfilter', dict(mode='clip', B=8, n_step=4, use_r=True, reward_scale=0.5, exp_interval=8), 2),
    ]
    for tag, kw, iters in variants:
        torch.manual_seed(11)
        lc, ec, sc = cfg_ppo(**kw)
        L = H.construct_without_initialize(PPOLearner, lc, ec, sc)
        L._ps_publisher = H._Any()
        L.tensorplex = H._Any()
        rng = np.random.default_rng(11)
        B, n, D, A = L.batch_size, L.n_step, 11, 3
        L.model.z_filter.z_update(torch.tensor((rng.standard_normal((40, D)) * 1.5 + 0.3).astype(np.float32)))
        L.ref_target_model.update_target_params(L.model)
        out = dict(**sd_np(L.model, 'init/'))
        n_pol = []
        orig_update = L._clip_update if L.ppo_mode == 'clip' else L._adapt_update

        def counting(*a, **k):
            n_pol[-1] += 1
            return orig_update(*a, **k)
        if L.ppo_mode == 'clip':
            L._clip_update = counting
        else:
            L._adapt_update = counting
        stats_all, hyper = [], []
        captured = {}
        orig_opt = L._optimize

        def opt_hook(*a, **k):
            st = orig_opt(*a, **k)
            captured['stats'] = {kk: float(vv) for kk, vv in st.items()}
            return st
        L._optimize = opt_hook
        for it in range(iters):
            batch, raw = make_ppo_batch(L, rng, B, n, D, A)
            for k, v in raw.items():
                out['it%d/%s' % (it, k)] = v
            n_pol.append(0)
            L.learn(batch)
            stats_all.append(captured['stats'])
            L.publish_parameter(it, message='')          # -> _post_publish when exp_counter >= exp_interval
            out.update(sd_np(L.model, 'it%d/after/' % it))
            out.update(sd_np(L.ref_target_model, 'it%d/ref/' % it))
            hyper.append(dict(clip_epsilon=getattr(L, 'clip_epsilon', None), beta=getattr(L, 'beta', None),
                              exp_counter=L.exp_counter, n_policy_epochs=n_pol[-1],
                              kl_record=list(map(float, L.kl_record))))
            if L.use_r_filter:
                out['it%d/rfilter' % it] = np.array([L.reward_filter.count.item(), L.reward_filter.running_sum.item(),
                                                     L.reward_filter.running_sumsq.item()])
        cfg = dict(mode=L.ppo_mode, B=B, n_step=n, D=D, A=A, actor_h=lc.model.actor_fc_hidden_sizes,
                   critic_h=lc.model.critic_fc_hidden_sizes, lr=lc.algo.network.lr_actor, gamma=L.gamma,
                   lam=L.lam, exp_interval=lc.parameter_publish.exp_interval, iters=iters,
                   use_r_filter=bool(L.use_r_filter), reward_scale=L.reward_scale,
                   kl_target=L.kl_target, epoch_policy=L.epoch_policy, epoch_baseline=L.epoch_baseline)
        save('ppo_learn_' + tag, cfg=cfg, stats=stats_all, hyper=hyper, **out)


def gen_ppo_learn_rnn():
    """RNN mode (the reference's DEFAULT PPO config: LSTM stem + horizon GAE, ppo.py:389-406,507-525): learn() +
    publish for a few iterations, LSTM trained by BOTH optimisers (ppo_net.py:202-224)."""
    variants = [
        ('rnn_clip', dict(mode='clip', B=8, n_step=7, rnn=True, horizon=3, rnn_hidden=10, exp_interval=16), 3),
        ('rnn_adapt', dict(mode='adapt', B=8, n_step=7, rnn=True, horizon=3, rnn_hidden=10, exp_interval=16), 3),
        ('rnn_adapt_biglr', dict(mode='adapt', B=8, n_step=6, rnn=True, horizon=2, rnn_hidden=6, lr=3e-3,
                                 exp_interval=8), 2),
    ]
    for tag, kw, iters in variants:
        torch.manual_seed(13)
        lc, ec, sc = cfg_ppo(**kw)
        L = H.construct_without_initialize(PPOLearner, lc, ec, sc)
        L._ps_publisher = H._Any()
        L.tensorplex = H._Any()
        rng = np.random.default_rng(13)
        B, n, D, A = L.batch_size, L.n_step, 11, 3
        L.model.z_filter.z_update(torch.tensor((rng.standard_normal((40, D)) * 1.5 + 0.3).astype(np.float32)))
        L.ref_target_model.update_target_params(L.model)
        out = dict(**sd_np(L.model, 'init/'))
        n_pol = []
        orig_update = L._clip_update if L.ppo_mode == 'clip' else L._adapt_update

        def counting(*a, **k):
            n_pol[-1] += 1
            return orig_update(*a, **k)
        if L.ppo_mode == 'clip':
            L._clip_update = counting
        else:
            L._adapt_update = counting
        stats_all, hyper = [], []
        captured = {}
        orig_opt = L._optimize
        orig_gae = L._gae_and_return

        def opt_hook(*a, **k):
            st = orig_opt(*a, **k)
            captured['stats'] = {kk: float(vv) for kk, vv in st.items()}
            return st

        def gae_hook(*a, **k):
            adv, ret = orig_gae(*a, **k)
            captured['adv'], captured['ret'] = adv.detach().clone().numpy(), ret.detach().clone().numpy()
            return adv, ret
        L._optimize = opt_hook
        L._gae_and_return = gae_hook
        for it in range(iters):
            batch, raw = make_ppo_batch(L, rng, B, n, D, A)
            for k, v in raw.items():
                out['it%d/%s' % (it, k)] = v
            n_pol.append(0)
            L.learn(batch)
            stats_all.append(captured['stats'])
            out['it%d/adv' % it], out['it%d/ret' % it] = captured['adv'], captured['ret']
            L.publish_parameter(it, message='')
            out.update(sd_np(L.model, 'it%d/after/' % it))
            out.update(sd_np(L.ref_target_model, 'it%d/ref/' % it))
            hyper.append(dict(clip_epsilon=getattr(L, 'clip_epsilon', None), beta=getattr(L, 'beta', None),
                              exp_counter=L.exp_counter, n_policy_epochs=n_pol[-1],
                              kl_record=list(map(float, L.kl_record))))
        cfg = dict(mode=L.ppo_mode, B=B, n_step=n, D=D, A=A, actor_h=lc.model.actor_fc_hidden_sizes,
                   critic_h=lc.model.critic_fc_hidden_sizes, lr=lc.algo.network.lr_actor, gamma=L.gamma,
                   lam=L.lam, exp_interval=lc.parameter_publish.exp_interval, iters=iters, horizon=L.horizon,
                   rnn_hidden=lc.algo.rnn.rnn_hidden, rnn_layer=lc.algo.rnn.rnn_layer,
                   kl_target=L.kl_target, epoch_policy=L.epoch_policy, epoch_baseline=L.epoch_baseline)
        save('ppo_learn_' + tag, cfg=cfg, stats=stats_all, hyper=hyper, **out)


def cfg_ddpg(D=9, A=3, actor_h=(20, 12), critic_h=(24, 16), B=16, n_step=3, target='hard', interval=2,
             tau=0.05, clip_critic=False, double=False):
    lc = Config(copy.deepcopy(DDPG_DEFAULT_LEARNER_CONFIG.to_dict()))
    ec = Config(copy.deepcopy(DDPG_DEFAULT_ENV_CONFIG.to_dict()))
    sc = Config(copy.deepcopy(DDPG_DEFAULT_SESSION_CONFIG.to_dict()))
    sc.folder = tempfile.mkdtemp()
    ec.env_name = 'synthetic'
    ec.num_agents = 4
    ec.obs_spec = {'low_dim': {'flat_inputs': (D,)}}
    ec.action_spec = {'dim': (A,), 'type': 'continuous'}
    ec.frame_stack_concatenate_on_env = True
    lc.model.actor_fc_hidden_sizes = list(actor_h)
    lc.model.critic_fc_hidden_sizes = list(critic_h)
    lc.algo.n_step = n_step
    lc.algo.network.clip_critic_gradient = clip_critic
    lc.algo.network.critic_gradient_value_clip = 0.01
    lc.algo.network.use_double_critic = double
    if target == 'hard':
        lc.algo.network.target_update = {'type': 'hard', 'interval': interval}
    else:
        lc.algo.network.target_update = {'type': 'soft', 'tau': tau}
    lc.replay.batch_size = B
    return lc, ec, sc


def gen_ddpg():
    for tag, kw in [('hard', dict()), ('soft_clipcritic', dict(target='soft', clip_critic=True))]:
        torch.manual_seed(21)
        lc, ec, sc = cfg_ddpg(**kw)
        L = H.construct_without_initialize(DDPGLearner, lc, ec, sc)
        L.tensorplex = H._Any()
        # make target != model at start so that the target path is really exercised
        with torch.no_grad():
            for p in L.model_target.parameters():
                p.add_(0.05 * torch.randn_like(p))
        rng = np.random.default_rng(21)
        B, D, A = L.batch_size, 9, 3
        out = dict(**sd_np(L.model, 'init/model/'), **sd_np(L.model_target, 'init/target/'))
        stats_all = []
        for it in range(3):
            raw = dict(obs=rng.standard_normal((B, D)).astype(np.float32),
                       obs_next=rng.standard_normal((B, D)).astype(np.float32),
                       actions=rng.uniform(-1, 1, (B, A)).astype(np.float32),
                       rewards=rng.standard_normal((B, 1)),                       # float64 (aggregator.py:101)
                       dones=(rng.random((B, 1)) < 0.2).astype(np.float64))
            batch = H._AttrDict(obs={'low_dim': {'flat_inputs': raw['obs'].copy()}},
                                obs_next={'low_dim': {'flat_inputs': raw['obs_next'].copy()}},
                                actions=raw['actions'].copy(), rewards=raw['rewards'].copy(),
                                dones=raw['dones'].copy())
            batch = L.preprocess(batch)
            st = L._optimize(batch.obs, batch.actions, batch.rewards, batch.obs_next, batch.dones)
            stats_all.append({k: float(v) for k, v in st.items() if not k.startswith('performance')})
            for k, v in raw.items():
                out['it%d/%s' % (it, k)] = v
            out.update(sd_np(L.model, 'it%d/model/' % it))
            out.update(sd_np(L.model_target, 'it%d/target/' % it))
        cfg = dict(B=B, D=D, A=A, actor_h=lc.model.actor_fc_hidden_sizes, critic_h=lc.model.critic_fc_hidden_sizes,
                   gamma=L.discount_factor, n_step=L.n_step, lr_actor=lc.algo.network.lr_actor,
                   lr_critic=lc.algo.network.lr_critic, target=lc.algo.network.target_update.to_dict(),
                   clip_actor=L.clip_actor_gradient, actor_clip=lc.algo.network.actor_gradient_value_clip,
                   clip_critic=L.clip_critic_gradient, critic_clip=lc.algo.network.critic_gradient_value_clip)
        save('ddpg_optimize_' + tag, cfg=cfg, stats=stats_all, **out)


def gen_ddpg_td3():
    """TD3 options of DDPGLearner._optimize (ddpg.py:267-283,298-321): second critic + target, y = min(y, y2) where
    only Q'_2 sees the (noised, clipped) target action -- Q'_1 was computed BEFORE the noise is added."""
    for tag, reg in (('td3_double', False), ('td3_double_reg', True)):
        torch.manual_seed(23)
        lc, ec, sc = cfg_ddpg(double=True, interval=2)
        lc.algo.network.use_action_regularization = reg
        L = H.construct_without_initialize(DDPGLearner, lc, ec, sc)
        L.tensorplex = H._Any()
        with torch.no_grad():
            for mdl in (L.model_target, L.model_target2):
                for p in mdl.parameters():
                    p.add_(0.05 * torch.randn_like(p))
        rng = np.random.default_rng(23)
        B, D, A = L.batch_size, 9, 3
        out = dict(**sd_np(L.model, 'init/model/'), **sd_np(L.model_target, 'init/target/'),
                   **sd_np(L.model2, 'init/model2/'), **sd_np(L.model_target2, 'init/target2/'))
        stats_all = []
        for it in range(3):
            raw = dict(obs=rng.standard_normal((B, D)).astype(np.float32),
                       obs_next=rng.standard_normal((B, D)).astype(np.float32),
                       actions=rng.uniform(-1, 1, (B, A)).astype(np.float32),
                       rewards=rng.standard_normal((B, 1)),
                       dones=(rng.random((B, 1)) < 0.2).astype(np.float64))
            batch = H._AttrDict(obs={'low_dim': {'flat_inputs': raw['obs'].copy()}},
                                obs_next={'low_dim': {'flat_inputs': raw['obs_next'].copy()}},
                                actions=raw['actions'].copy(), rewards=raw['rewards'].copy(),
                                dones=raw['dones'].copy())
            batch = L.preprocess(batch)
            np.random.seed(900 + it)
            raw['policy_noise_unclipped'] = np.random.normal(0, 0.2, size=(B, A))      # the draw _optimize will make
            np.random.seed(900 + it)
            st = L._optimize(batch.obs, batch.actions, batch.rewards, batch.obs_next, batch.dones)
            stats_all.append({k: float(v) for k, v in st.items() if not k.startswith('performance')})
            for k, v in raw.items():
                out['it%d/%s' % (it, k)] = v
            for name, mdl in (('model', L.model), ('target', L.model_target), ('model2', L.model2),
                              ('target2', L.model_target2)):
                out.update(sd_np(mdl, 'it%d/%s/' % (it, name)))
        cfg = dict(B=B, D=D, A=A, actor_h=lc.model.actor_fc_hidden_sizes, critic_h=lc.model.critic_fc_hidden_sizes,
                   gamma=L.discount_factor, n_step=L.n_step, lr_actor=lc.algo.network.lr_actor,
                   lr_critic=lc.algo.network.lr_critic, target=lc.algo.network.target_update.to_dict(),
                   clip_actor=L.clip_actor_gradient, actor_clip=lc.algo.network.actor_gradient_value_clip,
                   clip_critic=L.clip_critic_gradient, critic_clip=lc.algo.network.critic_gradient_value_clip,
                   action_regularization=reg)
        save('ddpg_optimize_' + tag, cfg=cfg, stats=stats_all, **out)


def gen_replay():
    # FIFO: ids in, ids out; capacity memory_size + 3 silently drops oldest (fifo_replay.py:27)
    lc, ec, sc = cfg_ppo()
    lc.replay.batch_size = 4
    lc.replay.memory_size = 6
    R = H.construct_without_initialize(FIFOReplay, lc, ec, sc)     # plain class, no metaclass
    script, trace = [], []
    nxt = 0
    rng = np.random.default_rng(5)
    for _ in range(60):
        if rng.random() < 0.65:
            k = int(rng.integers(1, 5))
            for _ in range(k):
                R.insert({'id': nxt})
                nxt += 1
            script.append(['insert', k])
            trace.append([len(R), int(R.start_sample_condition())])
        elif R.start_sample_condition():
            got = [e['id'] for e in R.sample(4)]
            script.append(['sample', 4])
            trace.append(got)
    fifo = dict(memory_size=6, batch_size=4, script=script, trace=trace)

    lc, ec, sc = cfg_ddpg()
    lc.replay.memory_size = 37
    lc.replay.sampling_start_size = 5
    U = H.construct_without_initialize(UniformReplay, lc, ec, sc)
    random.seed(5)
    nxt = 0
    script, trace = [], []
    for step in range(40):
        k = int(rng.integers(1, 6))
        for _ in range(k):
            U.insert({'id': nxt})
            nxt += 1
        script.append(['insert', k])
        trace.append([len(U), int(U.start_sample_condition()), U._next_idx])
        if U.start_sample_condition():
            got = [e['id'] for e in U.sample(8)]
            script.append(['sample', 8])
            trace.append(got)
    uni = dict(memory_size=37, sampling_start_size=5, seed=5, script=script, trace=trace)

    # raw index streams of random.randint for several population sizes (Appendix A.6)
    streams = {}
    for m in [1, 2, 5, 64, 96, 3000, 333333, 10 ** 6, 2 ** 20, 2 ** 20 + 1]:
        random.seed(5)
        streams[str(m)] = [random.randint(0, m - 1) for _ in range(64)]
    random.seed(12345678901234567890)
    streams['bigseed_1000'] = [random.randint(0, 999) for _ in range(32)]
    save('replay', fifo=fifo, uniform=uni, streams=streams)


class _ScriptEnv:
    """Deterministic scripted env used only to drive the reference's wrappers."""
    metadata = {}

    def __init__(self, ep_lens, D=2):
        self.ep_lens = list(ep_lens)
        self.ep = -1
        self.t = 0
        self.g = 0          # global step id
        self.D = D

    def _ob(self):
        return {'low_dim': {'flat_inputs': np.full((self.D,), float(self.g), dtype=np.float32)}}

    def reset(self):
        self.ep += 1
        self.t = 0
        return self._ob(), {}

    def close(self):
        pass

    def step(self, action):
        self.t += 1
        self.g += 1
        reward = 0.25 * self.g - 3.0
        done = self.t >= self.ep_lens[self.ep]
        return self._ob(), reward, done, {}


class _Sink:
    def __init__(self):
        self.sent = []

    def send(self, hash_dict, nonhash_dict):
        self.sent.append((copy.deepcopy(hash_dict), copy.deepcopy(nonhash_dict)))


def gen_window():
    cases = []
    for (n, stride, ep_lens) in [(4, 4, [10, 3, 9]), (5, 2, [11, 4, 7]), (3, 5, [9, 8]), (25, 20, [60, 30])]:
        lc, ec, sc = cfg_ppo(n_step=n, stride=stride)
        env = _ScriptEnv(ep_lens)
        W = ESW.ExpSenderWrapperMultiStepMovingWindowWithInfo(env, lc, sc)
        W.sender = _Sink()
        for _ in ep_lens:
            ob, _ = W.reset()
            done = False
            while not done:
                g = env.g
                act = np.array([float(g)])
                info = [[], [np.array([float(g), 1.0])]]
                ob, r, done, _ = W.step((act, info))
        wins = []
        for hd, nd in W.sender.sent:
            wins.append(dict(obs=[int(o['low_dim']['flat_inputs'][0]) for o in hd['obs']],
                             obs_next=int(hd['obs_next']['low_dim']['flat_inputs'][0]),
                             actions=[float(a[0]) for a in nd['actions']],
                             rewards=[float(x) for x in nd['rewards']],
                             dones=[bool(x) for x in nd['dones']],
                             pd0=[float(p[0][0]) for p in nd['persistent_infos']],
                             n_step=nd['n_step']))
        cases.append(dict(n_step=n, stride=stride, ep_lens=ep_lens, windows=wins))
    save('window_multistep', cases=cases)

    cases = []
    for (n, gamma, ep_lens) in [(3, 0.99, [7, 2, 5]), (1, 0.9, [4, 3]), (5, 0.95, [12, 4])]:
        lc, ec, sc = cfg_ddpg(n_step=n)
        lc.algo.gamma = gamma
        env = _ScriptEnv(ep_lens)
        W = ESW.ExpSenderWrapperSSARNStepBootstrap(env, lc, sc)
        W.sender = _Sink()
        for _ in ep_lens:
            ob, _ = W.reset()
            done = False
            while not done:
                g = env.g
                ob, r, done, _ = W.step(np.array([float(g)]))
        recs = []
        for hd, nd in W.sender.sent:
            recs.append(dict(obs=int(hd['obs'][0]['low_dim']['flat_inputs'][0]),
                             obs_next=int(hd['obs'][1]['low_dim']['flat_inputs'][0]),
                             action=float(nd['action'][0]), reward=float(nd['reward']), done=bool(nd['done'])))
        cases.append(dict(n_step=n, gamma=gamma, ep_lens=ep_lens, records=recs))
    save('window_ssar', cases=cases)


def gen_aggregate():
    rng = np.random.default_rng(31)
    obs_spec = {'low_dim': {'flat_inputs': (4,)}}
    act_spec = {'dim': (2,), 'type': 'continuous'}
    B, n = 3, 5
    exps = []
    for b in range(B):
        exps.append(dict(
            obs=[{'low_dim': {'flat_inputs': rng.standard_normal(4).astype(np.float32)}} for _ in range(n)],
            obs_next={'low_dim': {'flat_inputs': rng.standard_normal(4).astype(np.float32)}},
            actions=[rng.standard_normal(2) for _ in range(n)],
            rewards=[float(rng.standard_normal()) for _ in range(n)],
            dones=[False] * (n - 1) + [bool(b % 2)],
            persistent_infos=[[rng.standard_normal(4).astype(np.float32)] for _ in range(n)],
            onetime_infos=[], infos=[{}] * n, n_step=n))
    agg = MultistepAggregatorWithInfo(obs_spec, act_spec).aggregate(exps)
    out = dict(
        ms_in_obs=np.stack([np.stack([o['low_dim']['flat_inputs'] for o in e['obs']]) for e in exps]),
        ms_in_obs_next=np.stack([e['obs_next']['low_dim']['flat_inputs'] for e in exps]),
        ms_in_actions=np.stack([np.stack(e['actions']) for e in exps]),
        ms_in_rewards=np.array([e['rewards'] for e in exps]),
        ms_in_dones=np.array([e['dones'] for e in exps]),
        ms_in_pd=np.stack([np.stack([p[0] for p in e['persistent_infos']]) for e in exps]),
        ms_obs=agg['obs']['low_dim']['flat_inputs'], ms_obs_next=agg['obs_next']['low_dim']['flat_inputs'],
        ms_actions=agg['actions'], ms_rewards=agg['rewards'], ms_dones=agg['dones'],
        ms_pd=agg['persistent_infos'][0], ms_onetime_is_none=int(agg['onetime_infos'] is None))
    dt = {k: str(v.dtype) for k, v in out.items() if isinstance(v, np.ndarray) and k.startswith('ms_') and not k.startswith('ms_in')}

    ss = []
    for b in range(4):
        ss.append(dict(obs=[{'low_dim': {'flat_inputs': rng.standard_normal(4).astype(np.float32)}},
                            {'low_dim': {'flat_inputs': rng.standard_normal(4).astype(np.float32)}}],
                       action=rng.uniform(-1, 1, 2), reward=float(rng.standard_normal()), done=bool(b == 2), info={}))
    a2 = SSARAggregator(obs_spec, act_spec).aggregate(ss)
    out.update(ss_in_obs=np.stack([e['obs'][0]['low_dim']['flat_inputs'] for e in ss]),
               ss_in_obs_next=np.stack([e['obs'][1]['low_dim']['flat_inputs'] for e in ss]),
               ss_in_action=np.stack([e['action'] for e in ss]),
               ss_in_reward=np.array([e['reward'] for e in ss]),
               ss_in_done=np.array([e['done'] for e in ss]),
               ss_obs=a2['obs']['low_dim']['flat_inputs'], ss_obs_next=a2['obs_next']['low_dim']['flat_inputs'],
               ss_actions=a2['actions'], ss_rewards=a2['rewards'], ss_dones=a2['dones'])
    dt.update({k: str(out[k].dtype) for k in ['ss_obs', 'ss_obs_next', 'ss_actions', 'ss_rewards', 'ss_dones']})
    save('aggregate', dtypes=dt, **out)


def gen_act():
    torch.manual_seed(41)
    lc, ec, sc = cfg_ppo(D=11, A=3)
    sc.agent.num_gpus = 0
    np.random.seed(41)
    Ag = H.construct_without_initialize(PPOAgent, lc, ec, sc, 3, 'training')
    noise = Ag.noise                                  # drawn in __init__ (ppo_agent.py:60-61)
    rng = np.random.default_rng(41)
    Ag.model.z_filter.z_update(torch.tensor((rng.standard_normal((40, 11)) * 1.5 + 0.3).astype(np.float32)))
    obs = rng.standard_normal((6, 11)).astype(np.float32)
    np.random.seed(4242)
    eps = np.random.randn(6, 1, 3)                    # the draws act() will make, in order
    np.random.seed(4242)
    acts, pds = [], []
    for i in range(6):
        a, info = Ag.act({'low_dim': {'flat_inputs': obs[i]}})
        acts.append(a)
        pds.append(info[1][0])
    Ag.agent_mode = 'eval_deterministic'
    det = np.stack([Ag.act({'low_dim': {'flat_inputs': obs[i]}}) for i in range(6)])
    save('ppo_act', obs=obs, eps=eps[:, 0, :], noise=noise, actions=np.stack(acts), pds=np.stack(pds),
         actions_det=det, action_dtype=str(acts[0].dtype), **sd_np(Ag.model, 'model/'))

    torch.manual_seed(42)
    lc, ec, sc = cfg_ddpg(D=9, A=3)
    ec.num_agents = 4
    Ag = H.construct_without_initialize(DDPGAgent, lc, ec, sc, 3, 'training')
    obs = rng.standard_normal((5, 9)).astype(np.float32)
    np.random.seed(777)
    eps = np.stack([np.random.normal(np.zeros(3), np.ones(3)) for _ in range(5)])   # unit draws, same stream
    np.random.seed(777)
    acts = np.stack([Ag.act({'low_dim': {'flat_inputs': obs[i]}}) for i in range(5)])
    save('ddpg_act', obs=obs, unit_noise=eps, sigma=Ag.sigma, actions=acts, **sd_np(Ag.model, 'model/'))


def gen_act_rnn():
    """PPOAgent.act in RNN mode (ppo_agent.py:84-93,133-137,169-183): the LSTM cells travel with the agent, the
    cells BEFORE each step are what the exp-sender ships as onetime_infos, reset() zeroes them."""
    torch.manual_seed(43)
    lc, ec, sc = cfg_ppo(D=11, A=3, rnn=True, horizon=3, rnn_hidden=10)
    sc.agent.num_gpus = 0
    np.random.seed(43)
    Ag = H.construct_without_initialize(PPOAgent, lc, ec, sc, 2, 'training')
    noise = Ag.noise
    rng = np.random.default_rng(43)
    Ag.model.z_filter.z_update(torch.tensor((rng.standard_normal((40, 11)) * 1.5 + 0.3).astype(np.float32)))
    obs = rng.standard_normal((7, 11)).astype(np.float32)
    np.random.seed(4343)
    eps = np.random.randn(7, 1, 3)
    np.random.seed(4343)
    acts, pds, h_before, c_before = [], [], [], []
    for i in range(7):
        if i == 4:
            Ag.reset()                                            # new episode: zero cells
        a, info = Ag.act({'low_dim': {'flat_inputs': obs[i]}})
        acts.append(a)
        pds.append(info[1][0])
        h_before.append(info[0][0])
        c_before.append(info[0][1])
    save('ppo_act_rnn', obs=obs, eps=eps[:, 0, :], noise=noise, actions=np.stack(acts), pds=np.stack(pds),
         h_before=np.stack(h_before), c_before=np.stack(c_before), reset_at=4, rnn_hidden=10,
         **sd_np(Ag.model, 'model/'))


def gen_ppo_learn_pixel():
    """Pixel mode (BASELINE cfg 4 shape, scaled down): uint8 frames -> /255 -> CNN stem (16k8s4, 32k4s2, FC) shared by
    actor and critic and trained by BOTH optimisers (ppo_net.py:136-140,202-224,268-273), no z-filter."""
    C, HW, F = 2, 20, 12
    for tag, mode in (('pixel_clip', 'clip'), ('pixel_adapt', 'adapt')):
        torch.manual_seed(17)
        lc, ec, sc = cfg_ppo(mode=mode, B=6, n_step=4, stride=4, use_z=False, exp_interval=12)
        ec.obs_spec = {'pixel': {'camera0': (C, HW, HW)}}
        ec.pixel_input = True
        lc.model.cnn_feature_dim = F
        L = H.construct_without_initialize(PPOLearner, lc, ec, sc)
        L._ps_publisher = H._Any()
        L.tensorplex = H._Any()
        L.ref_target_model.update_target_params(L.model)
        rng = np.random.default_rng(17)
        B, n, A = L.batch_size, L.n_step, 3
        out = dict(**sd_np(L.model, 'init/'))
        stats_all, hyper, captured = [], [], {}
        orig_opt, orig_gae = L._optimize, L._gae_and_return

        def opt_hook(*a, **k):
            st = orig_opt(*a, **k)
            captured['stats'] = {kk: float(vv) for kk, vv in st.items()}
            return st

        def gae_hook(*a, **k):
            adv, ret = orig_gae(*a, **k)
            captured['adv'], captured['ret'] = adv.detach().clone().numpy(), ret.detach().clone().numpy()
            return adv, ret
        L._optimize, L._gae_and_return = opt_hook, gae_hook
        for it in range(2):
            obs = rng.integers(0, 256, size=(B, n, C, HW, HW), dtype=np.uint8)
            obs_next = rng.integers(0, 256, size=(B, 1, C, HW, HW), dtype=np.uint8)
            with torch.no_grad():
                flat = {'pixel': {'camera0': torch.tensor(obs.reshape(B * n, C, HW, HW), dtype=torch.float32)}}
                pd = L.model.forward_actor(flat).numpy().reshape(B, n, 2 * A)
            pd = pd.copy()
            pd[:, :, A:] *= np.exp(rng.uniform(-0.25, 0.25, size=(B, 1, 1))).astype(np.float32)
            actions = np.clip(rng.standard_normal((B, n, A)) * pd[:, :, A:] + pd[:, :, :A], -1, 1)
            rewards = rng.standard_normal((B, n)) * 0.5 + 0.1
            dones = np.zeros((B, n), dtype=np.float32)
            dones[rng.random(B) < 0.3, n - 1] = 1.0
            batch = H._AttrDict(obs={'pixel': {'camera0': obs.copy()}}, obs_next={'pixel': {'camera0': obs_next.copy()}},
                                actions=actions.copy(), rewards=rewards.copy(), dones=dones.copy(),
                                persistent_infos=[pd.astype(np.float32).copy()], onetime_infos=None)
            for k, v in dict(obs=obs, obs_next=obs_next, actions=actions, rewards=rewards, dones=dones,
                             pd=pd.astype(np.float32)).items():
                out['it%d/%s' % (it, k)] = v
            L.learn(batch)
            stats_all.append(captured['stats'])
            out['it%d/adv' % it], out['it%d/ret' % it] = captured['adv'], captured['ret']
            L.publish_parameter(it, message='')
            out.update(sd_np(L.model, 'it%d/after/' % it))
            hyper.append(dict(clip_epsilon=getattr(L, 'clip_epsilon', None), beta=getattr(L, 'beta', None),
                              exp_counter=L.exp_counter, kl_record=list(map(float, L.kl_record))))
        cfg = dict(mode=L.ppo_mode, B=B, n_step=n, A=A, C=C, HW=HW, cnn_feature_dim=F,
                   actor_h=lc.model.actor_fc_hidden_sizes, critic_h=lc.model.critic_fc_hidden_sizes,
                   lr=lc.algo.network.lr_actor, exp_interval=lc.parameter_publish.exp_interval, iters=2)
        save('ppo_learn_' + tag, cfg=cfg, stats=stats_all, hyper=hyper, **out)


def gen_act_ou():
    """DDPGAgent.act with Ornstein-Uhlenbeck exploration (action_noise.py:22-39, ddpg_agent.py:128-134,176-183,205-208):
    float64 state, reset in pre_episode()."""
    torch.manual_seed(44)
    lc, ec, sc = cfg_ddpg(D=9, A=3)
    lc.algo.exploration.noise_type = 'ou_noise'
    lc.algo.exploration.theta = 0.15
    lc.algo.exploration.dt = 1e-3
    ec.num_agents = 4
    Ag = H.construct_without_initialize(DDPGAgent, lc, ec, sc, 3, 'training')
    rng = np.random.default_rng(44)
    obs = rng.standard_normal((8, 9)).astype(np.float32)
    np.random.seed(778)
    eps = np.stack([np.random.normal(size=3) for _ in range(8)])      # the unit draws __call__ will make, in order
    np.random.seed(778)
    acts, states = [], []
    for i in range(8):
        if i == 5:
            Ag.pre_episode()                                            # noise.reset()
        acts.append(Ag.act({'low_dim': {'flat_inputs': obs[i]}}))
        states.append(np.array(Ag.noise.x_prev, dtype=np.float64))
    save('ddpg_act_ou', obs=obs, unit_noise=eps, sigma=Ag.sigma, theta=0.15, dt=1e-3, reset_at=5,
         actions=np.stack(acts), ou_states=np.stack(states), **sd_np(Ag.model, 'model/'))


def gen_checkpoint():
    """Files written by the REFERENCE's PeriodicCheckpoint (utils/checkpoint.py:18-347) for a small tracked object:
    the raw bytes of every file in the folder, so that the product's loader can be tested against them."""
    from surreal.utils.checkpoint import PeriodicCheckpoint

    class Obj:
        pass
    torch.manual_seed(5)
    o = Obj()
    o.model = torch.nn.Linear(3, 2)
    o.optim = torch.optim.Adam(o.model.parameters(), lr=1e-3)
    o.current_iteration = 0
    folder = tempfile.mkdtemp()
    ck = PeriodicCheckpoint(folder, 'learner', period=2, min_interval=0, tracked_obj=o,
                            tracked_attrs=['model', 'optim', 'current_iteration'], keep_history=2, keep_best=1)
    weights = {}
    for step in range(1, 7):
        o.model.weight.data += 0.5
        o.current_iteration = step
        if ck.save(score=float(step % 4), global_steps=step):
            weights[str(step)] = o.model.weight.detach().numpy().copy()
    files = {fn: np.frombuffer(open(os.path.join(folder, fn), 'rb').read(), dtype=np.uint8)
             for fn in sorted(os.listdir(folder))}
    save('checkpoint', file_names=sorted(files), **{'file/' + k: v for k, v in files.items()},
         **{'weight/' + k: v for k, v in weights.items()})


def gen_ppo_learner_ckpt():
    """A checkpoint folder written by the REFERENCE for a real PPOLearner (ppo.py:668-678 attributes: model,
    ref_target_model, both LR schedulers, current_iteration; utils/checkpoint.py:234-314 format) after two learn() +
    publish iterations, plus the state the loader must end up with and one more batch with the reference's result of
    learning on it from the restored state with FRESH optimisers (the reference does not checkpoint Adam moments)."""
    from surreal.utils.checkpoint import PeriodicCheckpoint
    torch.manual_seed(21)
    lc, ec, sc = cfg_ppo(mode='clip', B=16, n_step=6, exp_interval=16)
    L = H.construct_without_initialize(PPOLearner, lc, ec, sc)
    L._ps_publisher = H._Any()
    L.tensorplex = H._Any()
    rng = np.random.default_rng(21)
    B, n, D, A = L.batch_size, L.n_step, 11, 3
    L.model.z_filter.z_update(torch.tensor((rng.standard_normal((40, D)) * 1.5 + 0.3).astype(np.float32)))
    L.ref_target_model.update_target_params(L.model)
    folder = tempfile.mkdtemp()
    ck = PeriodicCheckpoint(folder, 'learner', period=1, min_interval=0, tracked_obj=L,
                            tracked_attrs=L.checkpoint_attributes(), keep_history=2, keep_best=0)
    L.periodic_checkpoint = lambda **kw: ck.save(score=None, global_steps=kw.get('global_steps'))
    for it in range(2):
        batch, raw = make_ppo_batch(L, rng, B, n, D, A)
        L.learn(batch)
        L.publish_parameter(it, message='')
    ck.save(score=None, global_steps=L.current_iteration)          # state AFTER the second publish
    out = dict(**sd_np(L.model, 'saved/model/'), **sd_np(L.ref_target_model, 'saved/ref/'))
    files = {fn: np.frombuffer(open(os.path.join(folder, fn), 'rb').read(), dtype=np.uint8) for fn in sorted(os.listdir(folder))}
    # what the reference itself computes when it restores that folder into a fresh learner and learns one more batch
    torch.manual_seed(99)
    lc2, ec2, sc2 = cfg_ppo(mode='clip', B=16, n_step=6, exp_interval=16)
    L2 = H.construct_without_initialize(PPOLearner, lc2, ec2, sc2)
    L2._ps_publisher = H._Any()
    L2.tensorplex = H._Any()
    ck2 = PeriodicCheckpoint(folder, 'learner', period=1, min_interval=0, tracked_obj=L2,
                             tracked_attrs=L2.checkpoint_attributes(), keep_history=2, keep_best=0)
    assert ck2.restore(target=0, mode='history', check_ckpt_exists=True)
    L2.periodic_checkpoint = lambda **kw: None
    captured = {}
    orig_opt = L2._optimize

    def opt_hook(*a, **k):
        st = orig_opt(*a, **k)
        captured['stats'] = {kk: float(vv) for kk, vv in st.items()}
        return st
    L2._optimize = opt_hook
    batch, raw = make_ppo_batch(L2, rng, B, n, D, A)
    for k, v in raw.items():
        out['next/%s' % k] = v
    L2.learn(batch)
    out.update(sd_np(L2.model, 'next/after/'))
    save('ppo_learner_ckpt', file_names=sorted(files), cfg=dict(B=B, n_step=n, D=D, A=A, exp_interval=16,
         actor_h=lc.model.actor_fc_hidden_sizes, critic_h=lc.model.critic_fc_hidden_sizes, lr=lc.algo.network.lr_actor,
         current_iteration=int(L.current_iteration), clip_epsilon=float(L.clip_epsilon),
         restored_iteration=int(L2.current_iteration - 1), sched_n_step=int(L.actor_lr_scheduler.n_step)),
         next_stats=captured['stats'], **{'file/' + k: v for k, v in files.items()}, **out)


def gen_configs():
    """Default config trees exactly as the reference builds them (main/ppo_configs.py:15-175,
    main/ddpg_configs.py:16-174, session/default_configs.py:4-259)."""
    from surreal.session import BASE_LEARNER_CONFIG, BASE_ENV_CONFIG, BASE_SESSION_CONFIG, LOCAL_SESSION_CONFIG

    def js(c):
        return json.loads(json.dumps(c.to_dict() if hasattr(c, 'to_dict') else c, default=str))
    save('configs', ppo_learner=js(PPO_DEFAULT_LEARNER_CONFIG), ppo_env=js(PPO_DEFAULT_ENV_CONFIG),
         ppo_session=js(PPO_DEFAULT_SESSION_CONFIG), ddpg_learner=js(DDPG_DEFAULT_LEARNER_CONFIG),
         ddpg_env=js(DDPG_DEFAULT_ENV_CONFIG), ddpg_session=js(DDPG_DEFAULT_SESSION_CONFIG),
         base_learner=js(BASE_LEARNER_CONFIG), base_env=js(BASE_ENV_CONFIG), base_session=js(BASE_SESSION_CONFIG),
         local_session=js(dict(LOCAL_SESSION_CONFIG)))


if __name__ == '__main__':
    which = sys.argv[1:] or ['pd', 'filters', 'gae', 'ppo_learn', 'ppo_learn_rnn', 'ppo_learn_pixel', 'ddpg', 'ddpg_td3',
                             'replay', 'window', 'aggregate', 'act', 'act_rnn', 'act_ou', 'checkpoint', 'ppo_learner_ckpt', 'configs']
    for w in which:
        globals()['gen_' + w]()
