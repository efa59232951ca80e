"""Harness that makes the *real* SurrealAI/surreal sources under /root/reference importable
in this container so golden vectors can be generated from the reference's own code.

The reference's third-party dependencies (benedict, torchx==0.9, caraml, tensorplex, symphony,
gym, mujoco ...) are absent here and cannot be installed (no network).  None of them holds
arithmetic of the hot path: they are process plumbing, logging, and thin
lazy-shape wrappers around torch.nn layers.  This module registers minimal stand-ins for them
in ``sys.modules`` -- nothing under /root/reference is modified or copied.

The torchx layer stand-ins map 1:1 onto stock torch.nn layers (Linear / Conv2d / ReLU / Tanh /
LayerNorm / Flatten); golden fixtures always INJECT weights, so torchx's unknown default
initialisation never enters a fixture.

Only used by ``make_golden.py`` (run in the build container).  Never imported by the product,
by tests at run time, or on the GPU box (where /root/reference does not exist).
"""
import collections
import collections.abc
import contextlib
import os
import sys
import types

import numpy as np
import torch.nn as nn

REF_ROOT = '/root/reference'


class _Any:
    """Absorbs any call / attribute access (logging + messaging sinks)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


class _MagicMod(types.ModuleType):
    __path__ = []          # behave as a package so `import x.y` resolves via sys.modules

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)

        class _C:
            def __init__(self, *a, **kw):
                pass
        _C.__name__ = k
        return _C


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---- torchx.layers stand-ins: lazily shaped layers over torch.nn -------------------------
class _Node:
    def __init__(self, shape, chain):
        self.shape = shape
        self.chain = chain


class _Lazy:
    def __call__(self, node):
        m, out_shape = self.make(node.shape)
        return _Node(out_shape, node.chain + [m])


class Linear(_Lazy):
    def __init__(self, out):
        self.out = out

    def make(self, s):
        return nn.Linear(s[-1], self.out), (*s[:-1], self.out)


class Conv2d(_Lazy):
    def __init__(self, out, kernel_size, stride=1):
        self.out, self.k, self.s = out, kernel_size, stride

    def make(self, s):
        c, h, w = s[-3:]
        ho = (h - self.k) // self.s + 1
        wo = (w - self.k) // self.s + 1
        return nn.Conv2d(c, self.out, self.k, self.s), (*s[:-3], self.out, ho, wo)


class ReLU(_Lazy):
    def make(self, s):
        return nn.ReLU(), s


class Tanh(_Lazy):
    def make(self, s):
        return nn.Tanh(), s


class Flatten(_Lazy):
    def make(self, s):
        return nn.Flatten(), (s[0], int(np.prod(s[1:])))


class LayerNorm(_Lazy):
    def __init__(self, n):
        self.n = n

    def make(self, s):
        return nn.LayerNorm(s[-self.n:]), s


def Placeholder(shape):
    return _Node(tuple(shape), [])


class Functional(nn.Module):
    def __init__(self, inputs, outputs):
        super().__init__()
        self.seq = nn.Sequential(*outputs.chain)

    def build(self, shape):
        pass

    def forward(self, x):
        return self.seq(x)


class Sequential(nn.Module):
    def __init__(self, *layers):
        super().__init__()
        self._lazy = layers

    def build(self, shape):
        node = _Node(tuple(shape), [])
        for layer in self._lazy:
            node = layer(node)
        self.seq = nn.Sequential(*node.chain)

    def forward(self, x):
        return self.seq(x)


class _TxModule(nn.Module):
    """torchx.nn.Module helpers used by surreal/learner/ddpg.py:174-178,309,332,410-428."""

    def hard_update(self, other):
        self.load_state_dict(other.state_dict())

    def soft_update(self, other, tau):
        for p, q in zip(self.parameters(), other.parameters()):
            p.data.mul_(1.0 - tau).add_(q.data, alpha=tau)

    def clip_grad_value(self, value):
        nn.utils.clip_grad_value_(self.parameters(), value)


class LinearWithMinLR:
    """Stand-in for torchx.nn.hyper_scheduler.LinearWithMinLR (source absent).
    Fixtures keep the LR constant, so the schedule formula never enters a golden vector."""

    def __init__(self, optim, num_updates, update_freq=1, min_lr=0):
        self.optim = optim
        self.n_step = 0

    def step(self):
        self.n_step += 1

    def get_lr(self):
        return [g['lr'] for g in self.optim.param_groups]

    def state_dict(self):
        return {'n_step': self.n_step}


_installed = False


def install():
    """Register the stand-ins and put /root/reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    _installed = True
    for n in ('Sequence', 'Mapping', 'Iterable', 'Callable', 'MutableMapping'):
        if not hasattr(collections, n):        # py3.10+ removed the aliases (utils/common.py:137)
            setattr(collections, n, getattr(collections.abc, n))
    _mod('benedict', BeneDict=_AttrDict)
    _mod('tensorplex', TensorplexClient=_Any, LoggerplexClient=_Any, Tensorplex=_Any, Loggerplex=_Any)
    cz = _mod('caraml')
    czz = _mod('caraml.zmq', **{k: _Any for k in [
        'ZmqServer', 'ZmqClient', 'ZmqPub', 'ZmqSub', 'ZmqSender', 'ZmqReceiver',
        'ZmqProxyThread', 'DataFetcher', 'ZmqPusher', 'ZmqPuller', 'ZmqTimeoutError',
        'get_remote_client']})
    cz.zmq = czz
    txnn = _mod('torchx.nn', Module=_TxModule)
    hs = _mod('torchx.nn.hyper_scheduler', LinearWithMinLR=LinearWithMinLR)
    hs.__all__ = ['LinearWithMinLR']
    txl = _mod('torchx.layers', Linear=Linear, Conv2d=Conv2d, ReLU=ReLU, Tanh=Tanh,
               Flatten=Flatten, LayerNorm=LayerNorm, Placeholder=Placeholder,
               Functional=Functional, Sequential=Sequential)
    tx = _mod('torchx', nn=txnn, layers=txl,
              device_scope=lambda *a, **k: contextlib.nullcontext())
    txnn.hyper_scheduler = hs
    _mod('pkg_resources',
         parse_version=lambda v: tuple(int(x) for x in v.split('+')[0].split('.')[:3]))
    for n in ['gym', 'gym.spaces', 'gym.wrappers', 'cv2', 'dm_control', 'dm_control.suite',
              'dm_control.rl', 'dm_control.rl.environment', 'dm_control.suite.wrappers',
              'dm_control.suite.wrappers.pixels', 'robosuite', 'robosuite.wrappers', 'mujoco_py',
              'imageio', 'tensorboardX', 'nanolog', 'symphony', 'symphony.engine',
              'symphony.commandline', 'symphony.addons', 'cloudwise', 'docker', 'pygame', 'PIL']:
        sys.modules[n] = _MagicMod(n)
    for v in ['PS_FRONTEND', 'PS_BACKEND', 'COLLECTOR_FRONTEND', 'COLLECTOR_BACKEND',
              'SAMPLER_FRONTEND', 'SAMPLER_BACKEND', 'PARAMETER_PUBLISH', 'PREFETCH_QUEUE',
              'TENSORPLEX', 'LOGGERPLEX']:
        os.environ.setdefault('SYMPH_%s_HOST' % v, '127.0.0.1')
        os.environ.setdefault('SYMPH_%s_PORT' % v, '7000')
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


class _AttrDict(dict):
    """benedict.BeneDict stand-in: dict with attribute access (learner.learn uses batch.obs)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    # BeneDict's YAML helpers (used by surreal/utils/checkpoint.py for the metadata file).  benedict is absent, so
    # the exact dump STYLE is unpinned (block style assumed); any YAML loader reads either style, which is all the
    # cross-implementation checkpoint test needs.
    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, _AttrDict) else v) for k, v in self.items()}

    def dump_yaml_file(self, path):
        import yaml
        with open(os.path.expanduser(path), 'w') as fp:
            yaml.safe_dump(self.to_dict(), fp, default_flow_style=False)

    @classmethod
    def load_yaml_file(cls, path):
        import yaml
        with open(os.path.expanduser(path)) as fp:
            d = yaml.safe_load(fp)
        return cls({k: (cls(v) if isinstance(v, dict) else v) for k, v in d.items()})


def construct_without_initialize(cls, *args, **kwargs):
    """Run cls.__init__ but skip AutoInitializeMeta._initialize (surreal/utils/common.py:270-275),
    which only wires ZeroMQ publishers / prefetch processes / logging threads."""
    obj = cls.__new__(cls)
    cls.__init__(obj, *args, **kwargs)
    return obj
