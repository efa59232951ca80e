"""Shared builders for tests: config trees shaped like the golden fixtures."""
import copy
import tempfile

import numpy as np
import torch


def ppo_configs(D=11, A=3, actor_h=(32, 24), critic_h=(28, 20), n_step=6, stride=6, B=16, mode='clip', lr=1e-4,
                exp_interval=4096, use_z=True, use_r=False, reward_scale=1.0, norm_adv=True, memory_size=None):
    from surreal_b200.session import Config
    from surreal_b200.main.ppo_configs import (PPO_DEFAULT_LEARNER_CONFIG, PPO_DEFAULT_ENV_CONFIG,
                                               PPO_DEFAULT_SESSION_CONFIG)
    lc = Config(copy.deepcopy(PPO_DEFAULT_LEARNER_CONFIG.to_dict()))
    ec = Config(copy.deepcopy(PPO_DEFAULT_ENV_CONFIG.to_dict()))
    sc = Config(copy.deepcopy(PPO_DEFAULT_SESSION_CONFIG.to_dict()))
    sc.folder = tempfile.mkdtemp()
    ec.obs_spec = {'low_dim': {'flat_inputs': (D,)}}
    ec.action_spec = {'dim': (A,), 'type': 'continuous'}
    lc.model.actor_fc_hidden_sizes = list(actor_h)
    lc.model.critic_fc_hidden_sizes = list(critic_h)
    lc.algo.ppo_mode = mode
    lc.algo.rnn.if_rnn_policy = False
    lc.algo.n_step = n_step
    lc.algo.stride = stride
    lc.algo.network.lr_actor = lr
    lc.algo.network.lr_critic = lr
    lc.algo.use_z_filter = use_z
    lc.algo.use_r_filter = use_r
    lc.algo.advantage.reward_scale = reward_scale
    lc.algo.advantage.norm_adv = norm_adv
    lc.replay.batch_size = B
    if memory_size is not None:
        lc.replay.memory_size = memory_size
    lc.parameter_publish.exp_interval = exp_interval
    return lc, ec, sc


def ddpg_configs(D=9, A=3, actor_h=(20, 12), critic_h=(24, 16), B=16, n_step=3, target=None, clip_critic=False,
                 lr_actor=1e-4, lr_critic=1e-3, memory_size=None, start=None):
    from surreal_b200.session import Config
    from surreal_b200.main.ddpg_configs import (DDPG_DEFAULT_LEARNER_CONFIG, DDPG_DEFAULT_ENV_CONFIG,
                                                DDPG_DEFAULT_SESSION_CONFIG)
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
    lc.algo.network.lr_actor = lr_actor
    lc.algo.network.lr_critic = lr_critic
    lc.algo.network.clip_critic_gradient = clip_critic
    lc.algo.network.critic_gradient_value_clip = 0.01
    if target is not None:
        lc.algo.network.target_update = target
    lc.replay.batch_size = B
    if memory_size is not None:
        lc.replay.memory_size = memory_size
    if start is not None:
        lc.replay.sampling_start_size = start
    return lc, ec, sc


def ref_state_dict(sub):
    """golden 'a/b/c' keys -> reference-style 'a.b.c' state_dict of torch tensors."""
    return {k.replace('/', '.'): torch.tensor(np.asarray(v)) for k, v in sub.items()}


def ppo_batch(b):
    """golden per-iteration arrays -> aggregator-format batch dict (numpy, host)."""
    return {'obs': {'low_dim': {'flat_inputs': b['obs']}}, 'obs_next': {'low_dim': {'flat_inputs': b['obs_next']}},
            'actions': b['actions'], 'rewards': b['rewards'], 'dones': b['dones'], 'persistent_infos': [b['pd']],
            'onetime_infos': None}
