"""CPU-side tests of the product's HOST logic (no GPU): aggregators, CPython-exact index stream, trackers,
checkpoint format, launcher argument handling."""
import os
import pickle
import random

import numpy as np
import pytest
import yaml


def test_product_aggregators_match_reference(golden):
    from surreal_b200.learner.aggregator import MultistepAggregatorWithInfo, SSARAggregator
    g = golden('aggregate')
    dt = g.js('dtypes')
    obs_spec = {'low_dim': {'flat_inputs': (4,)}}
    act_spec = {'dim': (2,), 'type': 'continuous'}
    ob = lambda v: {'low_dim': {'flat_inputs': v}}  # noqa: E731
    B, n = g['ms_in_obs'].shape[:2]
    exps = [dict(obs=[ob(g['ms_in_obs'][b, k]) for k in range(n)], obs_next=ob(g['ms_in_obs_next'][b]),
                 actions=list(g['ms_in_actions'][b]), rewards=[float(x) for x in g['ms_in_rewards'][b]],
                 dones=[bool(x) for x in g['ms_in_dones'][b]],
                 persistent_infos=[[g['ms_in_pd'][b, k]] for k in range(n)], onetime_infos=[], infos=[{}] * n, n_step=n)
            for b in range(B)]
    out = MultistepAggregatorWithInfo(obs_spec, act_spec).aggregate(exps)
    got = dict(obs=out['obs']['low_dim']['flat_inputs'], obs_next=out['obs_next']['low_dim']['flat_inputs'],
               actions=out['actions'], rewards=out['rewards'], dones=out['dones'], pd=out['persistent_infos'][0])
    for k, v in got.items():
        np.testing.assert_array_equal(v, g['ms_' + k])
        assert str(v.dtype) == dt['ms_' + k]
    assert out['onetime_infos'] is None
    ss = [dict(obs=[ob(g['ss_in_obs'][b]), ob(g['ss_in_obs_next'][b])], action=g['ss_in_action'][b],
               reward=float(g['ss_in_reward'][b]), done=bool(g['ss_in_done'][b]), info={}) for b in range(4)]
    o2 = SSARAggregator(obs_spec, act_spec).aggregate(ss)
    got = dict(obs=o2['obs']['low_dim']['flat_inputs'], obs_next=o2['obs_next']['low_dim']['flat_inputs'],
               actions=o2['actions'], rewards=o2['rewards'], dones=o2['dones'])
    for k, v in got.items():
        np.testing.assert_array_equal(v, g['ss_' + k])
        assert str(v.dtype) == dt['ss_' + k]
    with pytest.raises(NotImplementedError):
        SSARAggregator(obs_spec, {'dim': (2,), 'type': 'discrete'})


def test_cpp_mt19937_reproduces_cpython_randint(golden):
    """The C++ generator inside libsurreal_b200 (host code) == random.randint, and keeps the global stream in step."""
    from surreal_b200.replay.uniform_replay import PyRandomStream
    streams = golden('replay').js('streams')
    for m, exp in streams.items():
        if m.startswith('bigseed'):
            rng, mm = random.Random(12345678901234567890), 1000
        else:
            rng, mm = random.Random(5), int(m)
        out = np.empty(len(exp), dtype=np.int64)
        PyRandomStream(rng).randint_fill(mm, len(exp), out)
        assert out.tolist() == exp
    random.seed(99)
    a = [random.randint(0, 332) for _ in range(10)] + [random.random()]
    random.seed(99)
    out = np.empty(10, dtype=np.int64)
    PyRandomStream().randint_fill(333, 10, out)
    assert out.tolist() + [random.random()] == a


def test_trackers_and_timers():
    from surreal_b200.utils import PeriodicTracker, MovingAverageRecorder, AutoInitializeMeta
    t = PeriodicTracker(3)
    assert [t.track_increment() for _ in range(7)] == [False, False, True, False, False, True, False]
    m = MovingAverageRecorder(0.5)
    assert m.add_value(2.0) == 2.0 and abs(m.add_value(4.0) - (2.0 * 0.5 + 4.0) / 1.5) < 1e-12

    class A(metaclass=AutoInitializeMeta):
        def __init__(self):
            self.order = ['init']

        def _initialize(self):
            self.order.append('initialize')

    class B(A):
        def __init__(self):
            super().__init__()
            self.order.append('sub-init')
    assert B().order == ['init', 'sub-init', 'initialize']      # _initialize runs AFTER the most-derived __init__


def test_checkpoint_format_and_roundtrip(tmp_path):
    from surreal_b200.checkpoint import PeriodicCheckpoint

    class Mod:
        def __init__(self, v):
            self.v = v

        def state_dict(self):
            return {'w': self.v}

        def load_state_dict(self, sd):
            self.v = sd['w']

    class Obj:
        pass
    o = Obj()
    o.model, o.current_iteration = Mod(1.5), 7
    ck = PeriodicCheckpoint(str(tmp_path), 'learner', period=2, min_interval=0, tracked_obj=o,
                            tracked_attrs=['model', 'current_iteration'], keep_history=2, keep_best=0)
    assert ck.save(global_steps=1) is False and ck.save(global_steps=2) is True
    o.model.v, o.current_iteration = 2.5, 9
    ck.save(global_steps=3)
    assert ck.save(global_steps=4) is True
    o.model.v, o.current_iteration = 3.5, 11
    ck.save(global_steps=5)
    ck.save(global_steps=6)
    meta = yaml.safe_load(open(tmp_path / 'metadata.learner.yml'))
    assert meta['history_ckpt_files'] == ['learner.6.ckpt', 'learner.4.ckpt'] and meta['save_counter'] == 3
    assert not os.path.exists(tmp_path / 'learner.2.ckpt')                      # keep_history = 2
    data = pickle.load(open(tmp_path / 'learner.4.ckpt', 'rb'))
    assert list(data.keys()) == ['model', 'current_iteration'] and data['model'] == {'w': 2.5}
    o2 = Obj()
    o2.model, o2.current_iteration = Mod(0.0), 0
    ck2 = PeriodicCheckpoint(str(tmp_path), 'learner', period=1, tracked_obj=o2, tracked_attrs=None)
    assert ck2.restore(target=1, mode='history').endswith('learner.4.ckpt')
    assert (o2.model.v, o2.current_iteration) == (2.5, 9)
    assert ck2.restore(target=5, mode='history') is None


def test_launcher_component_parsing(tmp_path):
    from surreal_b200.launch import Launcher

    class L(Launcher):
        def setup(self, args):
            self.got_args = args

        def launch(self, name):
            return name
    la = L()
    assert la.main(['learner', '--', '--env', 'synthetic', '--num-agents', '4']) == 'learner'
    assert la.got_args == ['--env', 'synthetic', '--num-agents', '4']
    from surreal_b200.main.ppo_configs import ppo_argparser
    a = ppo_argparser().parse_args(['--env', 'synthetic:64:8', '--num-agents', '1024', '--experiment-folder', str(tmp_path)])
    assert a.num_agents == 1024 and a.unit_test is False


def test_checkpoint_reads_and_extends_reference_written_folder(tmp_path, golden):
    """tests/golden/checkpoint.npz holds the raw files the REFERENCE's PeriodicCheckpoint wrote (pickled state_dicts of
    a torch module + optimiser + a plain attribute, metadata yml): our Checkpoint must restore from them (newest,
    older, best) and keep saving into the same folder with the same metadata schema."""
    import torch
    from surreal_b200.checkpoint import PeriodicCheckpoint
    g = golden('checkpoint')
    for fn in g.js('file_names'):
        (tmp_path / fn).write_bytes(bytes(g['file/' + fn]))
    ref_meta = yaml.safe_load(open(tmp_path / 'metadata.learner.yml'))

    class Obj:
        pass
    o = Obj()
    o.model = torch.nn.Linear(3, 2)
    o.optim = torch.optim.Adam(o.model.parameters(), lr=1e-3)
    o.current_iteration = -1
    ck = PeriodicCheckpoint(str(tmp_path), 'learner', period=1, tracked_obj=o, tracked_attrs=None)
    assert ck.restore(target=0, mode='history').endswith('learner.6.ckpt')
    assert o.current_iteration == 6 and np.array_equal(o.model.weight.detach().numpy(), g['weight/6'])
    assert ck.restore(target=1, mode='history').endswith('learner.4.ckpt')
    assert o.current_iteration == 4 and np.array_equal(o.model.weight.detach().numpy(), g['weight/4'])
    assert ck.restore(target=0, mode='best').endswith('learner.best-6.ckpt')
    assert o.current_iteration == 6
    assert ck.restore(target='6', mode='history').endswith('learner.6.ckpt')
    # keep writing where the reference stopped: counters continue, pruning follows keep_history / keep_best
    o.current_iteration = 7
    assert ck.save(score=3.0, global_steps=7) is True
    meta = yaml.safe_load(open(tmp_path / 'metadata.learner.yml'))
    assert set(meta.keys()) == set(ref_meta.keys())
    assert meta['save_counter'] == ref_meta['save_counter'] + 1 and meta['tracked_attrs'] == ref_meta['tracked_attrs']
    assert meta['history_ckpt_files'] == ['learner.7.ckpt', 'learner.6.ckpt']
    assert meta['best_ckpt_files'] == ['learner.best-7.ckpt'] and not os.path.exists(tmp_path / 'learner.best-6.ckpt')
    assert not os.path.exists(tmp_path / 'learner.4.ckpt')
    assert set(meta['ckpt']['learner.7.ckpt'].keys()) == set(ref_meta['ckpt']['learner.6.ckpt'].keys())
    data = pickle.load(open(tmp_path / 'learner.7.ckpt', 'rb'))
    assert list(data.keys()) == ['model', 'optim', 'current_iteration'] and set(data['optim'].keys()) == {'state', 'param_groups'}


def test_parameter_wire_semantics_on_cpu_tensors():
    """The collapsed parameter wire (surreal/distributed/parameter_server.py:20-303): a fetch with the version the
    client already holds returns nothing (the hash-cached REQ), a publish bumps the version and snapshots the weights
    -- later learner updates do NOT leak to actors before the next publish (lagged actors) -- and the snapshot keeps
    its buffers across publishes (stable addresses for graph replays)."""
    import torch
    from surreal_b200.distributed import ModuleDict, ParameterPublisher, ParameterClient, LocalHub

    class FlatModel:
        def __init__(self):
            self.w = torch.zeros(5)

        def flat_state(self):
            return {'w': self.w}

        def load_flat_state(self, st):
            self.w.copy_(st['w'])

    learner_m, actor_m = FlatModel(), FlatModel()
    pub = ParameterPublisher({'ppo': learner_m})
    cli = ParameterClient(pub)
    assert cli.fetch_parameter_with_info() == (None, None)          # nothing published yet
    learner_m.w += 1.0
    pub.publish(3, message='batch 3')
    state, info = cli.fetch_parameter_with_info()
    assert info['iteration'] == 3 and info['message'] == 'batch 3' and info['hash'] == pub.version == 1
    ModuleDict({'ppo': actor_m}).load(state)
    assert torch.equal(actor_m.w, torch.ones(5))
    snap_ptr = state['ppo']['__flat__']['w'].data_ptr()
    assert cli.fetch_parameter_with_info()[0] is None               # same version: cached
    learner_m.w += 1.0                                              # learner moves on; actors must not see it yet
    assert torch.equal(pub.fetch(None)[0]['ppo']['__flat__']['w'], torch.ones(5))
    pub.publish(4)
    state, info = cli.fetch_parameter_with_info()
    assert info['hash'] == 2 and torch.equal(state['ppo']['__flat__']['w'], torch.full((5,), 2.0))
    assert state['ppo']['__flat__']['w'].data_ptr() == snap_ptr      # persistent snapshot buffer
    with pytest.raises(AssertionError):
        ModuleDict({1: actor_m})
    LocalHub.reset()

    from surreal_b200.session import Config
    a, b = LocalHub.get(Config({'folder': '/tmp/x'})), LocalHub.get(Config({'folder': '/tmp/x'}))
    assert a is b and LocalHub.get(Config({'folder': '/tmp/y'})) is not a
    LocalHub.reset()


def test_lr_scheduler_linear_with_floor():
    """algo.network.anneal (ppo.py:121-125,171-178): linear decay refreshed every update_freq steps, floored at min_lr;
    the value lands in the optimiser's lr slot.  (torchx's exact formula is unpinned -- the source is absent.)"""
    from surreal_b200.learner.scheduler import make_lr_scheduler

    class Opt:
        lr = None

        def set_lr(self, v):
            self.lr = v
    o = Opt()
    s = make_lr_scheduler('LinearWithMinLR', o, 1e-3, num_updates=10, update_freq=2, min_lr=2e-4)
    seen = []
    for _ in range(12):
        s.step()
        seen.append(s.get_lr()[0])
    assert seen[0] == 1e-3 and seen[1] == pytest.approx(8e-4) and seen[2] == seen[1]          # refreshed every 2nd step
    assert seen[7] == pytest.approx(2e-4) and seen[-1] == 2e-4 and o.lr == 2e-4                # floor
    assert all(a >= b for a, b in zip(seen, seen[1:]))
    s2 = make_lr_scheduler('LinearWithMinLR', Opt(), 1e-3, 10, 2, 2e-4)
    s2.load_state_dict(s.state_dict())
    assert s2.get_lr() == s.get_lr() and s2.n_step == 12
    with pytest.raises(ValueError):
        make_lr_scheduler('Cosine', o, 1e-3, 10, 1, 0.0)


def test_lr_scheduler_restores_foreign_state_dicts():
    """A reference-written PPO learner checkpoint carries torchx's scheduler dict, not ours: torch-style keys
    are mapped onto the step count, unknown dicts leave the schedule untouched instead of raising KeyError."""
    import warnings
    from surreal_b200.learner.scheduler import LinearWithMinLR

    class Opt:
        lr = None

        def set_lr(self, v):
            self.lr = v

    o = Opt()
    s = LinearWithMinLR(o, 1e-3, num_updates=100, update_freq=1, min_lr=1e-5)
    for _ in range(7):
        s.step()
    own = s.state_dict()
    s2 = LinearWithMinLR(Opt(), 1e-3, 100, 1, 1e-5)
    s2.load_state_dict(own)
    assert s2.n_step == 7 and abs(s2.lr - s.lr) < 1e-15
    s3 = LinearWithMinLR(Opt(), 1e-3, 100, 1, 1e-5)
    s3.load_state_dict({'last_epoch': 7, 'base_lrs': [1e-3], '_step_count': 8})
    assert s3.n_step == 7 and abs(s3.lr - s.lr) < 1e-15 and abs(s3.optim.lr - s.lr) < 1e-15
    s4 = LinearWithMinLR(Opt(), 1e-3, 100, 1, 1e-5)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        s4.load_state_dict({'something_else': 3})
    assert s4.n_step == 0 and s4.lr == 1e-3 and len(w) == 1


def test_reference_written_ppo_learner_checkpoint_loads_without_its_classes(tmp_path, golden):
    """tests/golden/ppo_learner_ckpt.npz: the folder the REFERENCE's PeriodicCheckpoint wrote for a real PPOLearner.  It
    pickles the two LR schedulers as OBJECTS of a class that does not exist here (torchx in a real deployment,
    _ref_harness in the fixture): the loader must still restore models, counters and the schedulers' step counts."""
    import json
    import torch
    from surreal_b200.checkpoint import PeriodicCheckpoint
    from surreal_b200.learner.scheduler import LinearWithMinLR
    g = golden('ppo_learner_ckpt')
    cfg = g.js('cfg')
    for fn in g.js('file_names') if g['file_names'].ndim == 0 else [str(x) for x in g['file_names']]:
        (tmp_path / fn).write_bytes(bytes(g['file/' + fn]))

    class Mod:
        def __init__(self):
            self.sd = None

        def load_state_dict(self, sd):
            self.sd = sd

    class Opt:
        def set_lr(self, v):
            self.lr = v

    class Obj:
        pass
    o = Obj()
    o.model, o.ref_target_model = Mod(), Mod()
    o.actor_lr_scheduler = LinearWithMinLR(Opt(), 1e-4, 1000, 1, 1e-6)
    o.critic_lr_scheduler = LinearWithMinLR(Opt(), 1e-4, 1000, 1, 1e-6)
    o.current_iteration = 0
    ck = PeriodicCheckpoint(str(tmp_path), 'learner', period=1, tracked_obj=o, tracked_attrs=None)
    assert ck.restore(target=0, mode='history', check_ckpt_exists=True)
    assert o.current_iteration == cfg['current_iteration']
    assert o.actor_lr_scheduler.n_step == cfg['sched_n_step'] == o.critic_lr_scheduler.n_step
    saved = g.sub('saved/model/')
    assert set(k.replace('.', '/') for k in o.model.sd) == set(saved)
    for k, v in o.model.sd.items():
        np.testing.assert_array_equal(torch.as_tensor(v).numpy().reshape(-1), saved[k.replace('.', '/')].reshape(-1))
