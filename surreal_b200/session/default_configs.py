"""Base config trees every component is extended with (behavioural mirror of
surreal/session/default_configs.py:4-259; equality with the reference's trees is pinned by
tests/test_default_configs.py against tests/golden/configs.npz).

In this engine the host/port entries of the session tree are INERT: actors, replay and learner are
co-located on the GPU and the ZeroMQ wires collapse to device pointers.  They are kept so
that existing Surreal session configs load unchanged."""
from .config import extend_config

_REQ_S, _REQ_I, _REQ_B, _REQ_F = '_str_', '_int_', '_bool_', '_float_'

BASE_LEARNER_CONFIG = {
    'model': '_dict_',
    'algo': {
        'n_step': 1, 'gamma': _REQ_F, 'use_batchnorm': False, 'limit_training_episode_length': 0,
        'network': {'actor_regularization': 0.0, 'critic_regularization': 0.0},
    },
    'replay': {'batch_size': _REQ_I, 'replay_shards': 1},
    'parameter_publish': {'min_publish_interval': 0.3},
}

BASE_ENV_CONFIG = {
    'env_name': _REQ_S, 'sleep_time': 0.0,
    'video': {'record_video': False, 'max_videos': 10, 'record_every': 10, 'save_folder': None},
    'eval_mode': {}, 'action_spec': {}, 'obs_spec': {},
    'frame_stacks': 1, 'frame_stack_concatenate_on_env': True,
}


def _ckpt(**kw):
    d = {'restore_target': _REQ_I, 'mode': '_enum[best,history]_', 'keep_history': _REQ_I, 'keep_best': _REQ_I,
         'periodic': _REQ_I}
    d.update(kw)
    return d


BASE_SESSION_CONFIG = {
    'folder': _REQ_S,
    'replay': {
        'collector_frontend_host': _REQ_S, 'collector_frontend_port': _REQ_I,
        'collector_backend_host': _REQ_S, 'collector_backend_port': _REQ_I,
        'sampler_frontend_host': _REQ_S, 'sampler_frontend_port': _REQ_I,
        'sampler_backend_host': _REQ_S, 'sampler_backend_port': _REQ_I,
        'max_puller_queue': _REQ_I, 'evict_interval': _REQ_F, 'tensorboard_display': True,
    },
    'sender': {'flush_iteration': _REQ_I, 'flush_time': _REQ_I},
    'ps': {
        'parameter_serving_frontend_host': _REQ_S, 'parameter_serving_frontend_port': _REQ_I,
        'parameter_serving_backend_host': _REQ_S, 'parameter_serving_backend_port': _REQ_I,
        'shards': _REQ_I, 'publish_host': '_str', 'publish_port': _REQ_I,   # '_str' (sic) as in the reference
    },
    'tensorplex': {
        'host': _REQ_S, 'port': _REQ_I, 'tensorboard_port': _REQ_I, 'agent_bin_size': 8, 'max_processes': 4,
        'update_schedule': {'training_env': _REQ_I, 'eval_env': _REQ_I, 'eval_env_sleep': _REQ_I, 'agent': _REQ_I,
                            'learner': _REQ_I, 'learner_min_update_interval': _REQ_I},
    },
    'loggerplex': {
        'host': _REQ_S, 'port': _REQ_I, 'overwrite': False, 'level': 'info', 'show_level': True,
        'time_format': 'hms', 'enable_local_logger': _REQ_B, 'local_logger_level': 'info',
        'local_logger_time_format': 'hms',
    },
    'agent': {'fetch_parameter_mode': _REQ_S, 'fetch_parameter_interval': int},   # the type object, as upstream
    'learner': {'num_gpus': _REQ_I, 'prefetch_host': _REQ_S, 'prefetch_port': _REQ_I, 'prefetch_processes': _REQ_I,
                'max_prefetch_queue': _REQ_I, 'max_preprocess_queue': _REQ_I},
    'checkpoint': {
        'restore': _REQ_B, 'restore_folder': None,
        'learner': _ckpt(min_interval=_REQ_I),
        'agent': _ckpt(),
    },
}

_LOCAL = 'localhost'
LOCAL_SESSION_CONFIG = {
    'folder': _REQ_S,
    'replay': {
        'collector_frontend_host': _LOCAL, 'collector_frontend_port': 7001,
        'collector_backend_host': _LOCAL, 'collector_backend_port': 7002,
        'sampler_frontend_host': _LOCAL, 'sampler_frontend_port': 7003,
        'sampler_backend_host': _LOCAL, 'sampler_backend_port': 7004,
        'max_puller_queue': 10000, 'evict_interval': 0., 'tensorboard_display': True,
    },
    'sender': {'flush_iteration': _REQ_I, 'flush_time': 0},
    'ps': {
        'parameter_serving_frontend_host': _LOCAL, 'parameter_serving_frontend_port': 7005,
        'parameter_serving_backend_host': _LOCAL, 'parameter_serving_backend_port': 7006,
        'shards': 2, 'publish_host': _LOCAL, 'publish_port': 7007,
    },
    'tensorplex': {
        'host': _LOCAL, 'port': 7008, 'tensorboard_port': 6006,
        'update_schedule': {'training_env': 20, 'eval_env': 20, 'eval_env_sleep': 30, 'agent': 20, 'learner': 20,
                            'learner_min_update_interval': 30},
    },
    'loggerplex': {'host': _LOCAL, 'port': 7009, 'enable_local_logger': True},
    'agent': {'fetch_parameter_mode': 'episode', 'fetch_parameter_interval': 1},
    'learner': {'num_gpus': 0, 'prefetch_host': _LOCAL, 'prefetch_port': 7010, 'prefetch_processes': 2,
                'max_prefetch_queue': 10, 'max_preprocess_queue': 2},
    'checkpoint': {
        'restore': False, 'restore_folder': None,
        'learner': {'restore_target': 0, 'mode': 'history', 'keep_history': 2, 'keep_best': 0, 'periodic': 100000,
                    'min_interval': 15 * 60},
        'agent': {'restore_target': 0, 'mode': 'history', 'keep_history': 2, 'keep_best': 0, 'periodic': 100},
    },
}
LOCAL_SESSION_CONFIG = extend_config(LOCAL_SESSION_CONFIG, BASE_SESSION_CONFIG)
