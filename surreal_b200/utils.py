"""Small host-side helpers shared by the plugin classes (constructor protocol, timers, trackers).

AutoInitializeMeta mirrors the reference's constructor protocol (surreal/utils/common.py:232-275):
``_initialize()`` runs after the most-derived ``__init__``.  The timers keep the names the reference
reports under (learner/base.py:160-162, replay/base.py:139-141)."""
import logging
import threading
import time
from contextlib import contextmanager


class AutoInitializeMeta(type):
    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        if not hasattr(obj, '_initialize'):
            raise AssertionError('AutoInitializeMeta requires that subclass implements _initialize()')
        obj._initialize()
        return obj


class MovingAverageRecorder:
    """Exponential moving average with bias correction (first value seeds the average)."""

    def __init__(self, decay=0.95):
        self.decay = decay
        self.cum_value = 0.0
        self.normalization = 0.0
        self._lock = threading.Lock()

    def add_value(self, value):
        with self._lock:
            self.cum_value = self.cum_value * self.decay + value
            self.normalization = self.normalization * self.decay + 1.0
            return self.cum_value / self.normalization

    def cur_value(self):
        with self._lock:
            return self.cum_value / self.normalization if self.normalization > 0 else 0.0


class TimeRecorder:
    def __init__(self, decay=0.9995, max_seconds=10):
        self.moving_average = MovingAverageRecorder(decay)
        self.max_seconds = max_seconds
        self.started = False

    @contextmanager
    def time(self):
        t0 = time.time()
        yield None
        self.moving_average.add_value(min(self.max_seconds, time.time() - t0))

    def start(self):
        if self.started:
            raise RuntimeError('Starting a started timer')
        self.pre_time = time.time()
        self.started = True

    def lap(self):
        if not self.started:
            raise RuntimeError('Stopping a timer that is not started')
        now = time.time()
        self.moving_average.add_value(min(self.max_seconds, now - self.pre_time))
        self.pre_time = now

    def stop(self):
        self.lap()
        self.started = False

    @property
    def avg(self):
        return self.moving_average.cur_value()


class PeriodicTracker:
    """True once every ``period`` increments (surreal/session/tracker.py:10-44)."""

    def __init__(self, period, init_value=0, init_endpoint=0):
        assert isinstance(period, int) and period > 0
        self.period, self.value, self._endpoint = period, init_value, init_endpoint

    def _update(self):
        if self.value >= self._endpoint + self.period:
            self._endpoint += (self.value - self._endpoint) // self.period * self.period
            return True
        return False

    def track_increment(self, incr=1):
        self.value += incr
        return self._update()

    def track_absolute(self, value):
        self.value = value
        return self._update()


class TimedTracker:
    """True when at least ``interval`` seconds passed since the last True (utils/common.py:590-610)."""

    def __init__(self, interval):
        self.interval = interval
        self.last_time = time.time()

    def track_increment(self):
        now = time.time()
        if now - self.last_time >= self.interval:
            self.last_time = now
            return True
        return False


class ScalarSink:
    """Stand-in for the tensorplex client: keeps the last scalars per step (``add_scalars`` is the only
    call the hot path makes, learner/base.py:611; the ZMQ->TensorBoard multiplexer is out of scope)."""

    def __init__(self, name='learner'):
        self.name = name
        self.last = {}
        self.history = []
        self.keep_history = False

    def add_scalars(self, tag_to_scalar, global_step=None):
        self.last = dict(tag_to_scalar)
        if self.keep_history:
            self.history.append((global_step, dict(tag_to_scalar)))


def get_logger(name):
    log = logging.getLogger('surreal_b200.' + name)
    if not log.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter('[%(name)s] %(message)s'))
        log.addHandler(h)
        log.setLevel(logging.WARNING)
    return log


def obs_is_pixel(obs_spec):
    """True for a pixel-only observation spec ({'pixel': {'camera0': (C, H, W)}}, docs/env.md:79-108)."""
    return 'pixel' in obs_spec and len(obs_spec['pixel']) > 0


def obs_packed_dim(obs_spec):
    """Floats per observation in the HBM staging / replay records.  Low-dim observations: the concatenated feature width.
    Pixel observations: the uint8 frame is carried as an opaque run of C*H*W/4 32-bit words (only ever COPIED by the
    windowing / replay kernels, never computed on), so frames stay uint8 in HBM: 28 224 B per 4x84x84 frame."""
    if obs_is_pixel(obs_spec):
        if 'low_dim' in obs_spec and len(obs_spec['low_dim']) > 0:
            raise NotImplementedError('mixed low_dim + pixel observations are not supported by the HBM replay records')
        shape = tuple(obs_spec['pixel']['camera0'])
        n = 1
        for v in shape:
            n *= int(v)
        if n % 16 != 0:
            raise ValueError('pixel observations need C*H*W to be a multiple of 16 bytes, got %s' % (shape,))
        return n // 4
    return sum(v[0] for v in obs_spec['low_dim'].values())


def obs_flat(obs):
    """The per-actor rows [N, packed_dim] the staging kernels copy: the low-dim feature tensor, or the uint8 frames viewed
    as 32-bit words."""
    import torch
    if isinstance(obs, dict):
        if 'low_dim' in obs and len(obs['low_dim']) > 0:
            return obs['low_dim']['flat_inputs'] if 'flat_inputs' in obs['low_dim'] else next(iter(obs['low_dim'].values()))
        fr = obs['pixel']['camera0']
        return fr.reshape(fr.shape[0], -1).view(torch.float32)
    return obs


def record_obs_dim(learner_config, env_config):
    """Floats per observation row of the HBM staging / replay records and the learner's batch buffers: obs_packed_dim, plus
    -- for an RNN policy -- 2 * rnn_hidden trailing floats that carry the actor's LSTM cells (h | c) BEFORE it acted on that
    observation.  The reference ships the cells of a window's first step as ``onetime_infos``
    (ppo_agent.py:133-137, exp_sender_wrapper.py:244-264); here they simply ride with every observation row."""
    d = obs_packed_dim(env_config.obs_spec)
    rnn = learner_config.algo.rnn if 'rnn' in learner_config.algo else None
    if rnn is not None and rnn.if_rnn_policy:
        d += 2 * int(rnn.rnn_hidden)
    return d
