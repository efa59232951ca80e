"""The reference's parameter wire collapsed to device memory.

surreal/distributed/parameter_server.py:20-303 moves ``state_dict -> numpy -> pyarrow -> md5 -> ZMQ PUB ->
server shards -> REQ/REP`` between processes.  Actors and learner share one GPU here, so a publish is a
version bump plus a device-to-device copy into a snapshot that actors read; actors keep LAGGED weights
between their own fetches exactly like Surreal's (docs/ppo.md:16, agent/base.py:182-189)."""
import time

import torch


class ModuleDict:
    """name -> model with state_dict()/load_state_dict() (surreal/distributed/module_dict.py:8-63)."""

    def __init__(self, module_dict):
        assert isinstance(module_dict, dict)
        for k in module_dict:
            assert isinstance(k, str), 'Key "{}" must be string.'.format(k)
        self._module_dict = module_dict

    def items(self):
        return self._module_dict.items()

    def dumps(self):
        return {k: {n: v.detach().cpu().numpy() for n, v in m.state_dict().items()}
                for k, m in self._module_dict.items()}

    def load(self, state):
        for k, m in self._module_dict.items():
            if '__flat__' in state[k]:
                m.load_flat_state(state[k]['__flat__'])
            else:
                m.load_state_dict(state[k])


class ParameterPublisher:
    """publish(iteration) snapshots every module's flat device buffers; ``version`` plays the role of the
    md5 content hash (parameter_server.py:40-55)."""

    def __init__(self, module_dict):
        if not isinstance(module_dict, ModuleDict):
            module_dict = ModuleDict(module_dict)
        self._module_dict = module_dict
        self.version = 0
        self.info = None
        self._snapshot = None

    def publish(self, iteration, message=''):
        snap = {}
        for name, m in self._module_dict.items():
            if hasattr(m, 'flat_state'):       # flat device buffers: a publish is a handful of D2D copies
                old = (self._snapshot or {}).get(name, {}).get('__flat__')
                cur = m.flat_state()
                if old is not None and old.keys() == cur.keys() and all(old[k].shape == cur[k].shape for k in cur):
                    for k, v in cur.items():       # persistent snapshot buffers: stable addresses, no allocator
                        old[k].copy_(v)            # traffic; cross-stream ordering is the engine's job (events)
                    snap[name] = {'__flat__': old}
                else:
                    snap[name] = {'__flat__': {k: v.detach().clone() for k, v in cur.items()}}
            else:
                snap[name] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        self._snapshot = snap
        self.version += 1
        self.info = {'time': time.time(), 'iteration': iteration, 'message': message, 'hash': self.version}

    def fetch(self, last_version):
        """-> (state or None, info): None when the caller already holds ``last_version``
        (the hash-cached REQ of parameter_server.py:200-204,241-262)."""
        if self._snapshot is None:
            return None, None
        if last_version == self.version:
            return None, self.info
        return self._snapshot, self.info


class ParameterClient:
    def __init__(self, publisher):
        self._publisher = publisher
        self._last_version = None

    def attach(self, publisher):
        self._publisher = publisher

    def fetch_parameter_with_info(self):
        if self._publisher is None:
            return None, None
        state, info = self._publisher.fetch(self._last_version)
        if state is not None:
            self._last_version = info['hash']
        return state, info

    def fetch_info(self):
        return self._publisher.info if self._publisher is not None else None


class LocalHub:
    """In-process rendezvous that replaces the SYMPH_* host/port environment variables
    (launch/setup_network.py:21-45): components of one experiment (same session folder) find each other
    here -- replay shards, the learner's parameter publisher -- instead of dialling ZeroMQ sockets."""
    _hubs = {}

    def __init__(self):
        self.replays = {}
        self.publisher = None

    @classmethod
    def get(cls, session_config):
        key = session_config.folder if 'folder' in session_config else '__default__'
        if key not in cls._hubs:
            cls._hubs[key] = LocalHub()
        return cls._hubs[key]

    @classmethod
    def reset(cls):
        cls._hubs = {}
