"""Replay base class: same constructor and method surface as surreal/replay/base.py:9-256.  The collector
and sampler ZeroMQ servers are gone -- actors write into the HBM store through kernels, the learner
pulls batches in-process -- so start_threads()/join() are no-ops kept for launcher compatibility."""
import time

from .. import utils as U
from ..distributed import LocalHub


def gather_fields(pairs, idx32, idx64, batch):
    """ONE launch for every field of the sampled records: ``pairs`` = [(src ring, out buffer, floats per record)]."""
    import ctypes as C
    from .. import _lib
    from ..ops import _stream
    n = len(pairs)
    srcs = (C.c_void_p * n)(*[p[0].data_ptr() for p in pairs])
    outs = (C.c_void_p * n)(*[p[1].data_ptr() for p in pairs])
    recs = (C.c_int64 * n)(*[int(p[2]) for p in pairs])
    _lib.check(_lib.lib().sb200_replay_gather_multi_f32(
        srcs, outs, recs, n, C.c_void_p(idx32.data_ptr()) if idx32 is not None else None,
        C.c_void_p(idx64.data_ptr()) if idx64 is not None else None, int(batch), _stream()), 'sb200_replay_gather_multi_f32')


class Replay:
    def __init__(self, learner_config, env_config, session_config, index=0):
        self.learner_config = learner_config
        self.env_config = env_config
        self.session_config = session_config
        self.index = index
        self.log = U.get_logger('replay/%d' % index)
        self.tensorplex = U.ScalarSink('replay/%d' % index)
        self.init_time = time.time()
        self.cumulative_collected_count = 0
        self.cumulative_sampled_count = 0
        self.cumulative_request_count = 0
        self.insert_time = U.TimeRecorder(decay=0.99998)
        self.sample_time = U.TimeRecorder()
        LocalHub.get(session_config).replays[index] = self

    def start_threads(self):
        pass

    def join(self):
        pass

    def insert(self, exp_dict):
        raise NotImplementedError

    def sample(self, batch_size):
        raise NotImplementedError

    def evict(self):
        pass

    def start_sample_condition(self):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def _insert_wrapper(self, exp):
        self.cumulative_collected_count += 1
        with self.insert_time.time():
            self.insert(exp)

    def sample_request(self, batch_size):
        """What the reference's _sample_request_handler does (base.py:156-171) minus serialisation."""
        while not self.start_sample_condition():
            time.sleep(0.01)
        self.cumulative_sampled_count += batch_size
        self.cumulative_request_count += 1
        with self.sample_time.time():
            return self.sample(batch_size)
