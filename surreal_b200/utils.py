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
