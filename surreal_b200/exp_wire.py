"""Experience ingress from EXTERNAL CPU actors: the message structure and transport of the reference's
experience wire -- surreal/distributed/exp_sender.py:10-98 (ExpBuffer / ExpSender), exp_collector.py:7-65
(ExperienceCollectorServer), utils/serializer.py:11-70 -- so that real MuJoCo / robosuite actor processes, which cannot
live on the GPU, can feed the HBM replay of a surreal_b200 learner.

  message  = serialize((exp_list, ob_storage)); every heavy value under a ``hash_dict`` key travels ONCE per message in
             ``ob_storage`` keyed by a 16-character content hash (md5 -> base64), the experience dicts carry
             ``<key>_hash`` references (overlapping windows ship each observation once); the collector re-inflates them and
             keeps a weak cache across messages.
  transport = ZeroMQ PUSH (actors) -> PULL (collector), one frame per flush (caraml.zmq.ZmqSender / ZmqReceiver).

Serializer: the reference defaults to ``pyarrow.serialize`` -- pyarrow's legacy format, REMOVED from pyarrow since 15.0 (this
image ships 24.0: ``pa.serialize`` does not exist), so it cannot be produced or parsed here.  The reference makes the
serializer pluggable (``surreal.utils.serializer.set_global_serializer``, serializer.py:24-32, with the pickle pair spelled
out at :20-21): actors that feed this collector call ``U.set_global_serializer(pickle.dumps, pickle.loads)``.  When a
pyarrow that still has the legacy format is importable it is used automatically for frames that are not pickles.
"""
import base64
import hashlib
import pickle
import threading
import weakref


def _legacy_pyarrow():
    try:
        import pyarrow as pa
        if hasattr(pa, 'serialize') and hasattr(pa, 'deserialize'):
            return pa
    except Exception:                                             # noqa: BLE001
        pass
    return None


def serialize(obj):
    return pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)


def deserialize(binary):
    if bytes(binary[:1]) == b'\x80':                               # pickle protocol >= 2 frame
        return pickle.loads(binary)
    pa = _legacy_pyarrow()
    if pa is None:
        raise ValueError('frame is not a pickle and this pyarrow has no legacy (de)serializer: have the actors call '
                         'surreal.utils.set_global_serializer(pickle.dumps, pickle.loads)')
    return pa.deserialize(binary)


def binary_hash(binary):
    """serializer.py:55-67: 16 characters of base64(md5)."""
    return base64.b64encode(hashlib.md5(binary).digest())[:16].decode('utf-8')


def pyobj_hash(obj):
    return binary_hash(serialize(obj))


class _Leaf:
    """Wrapper that makes numpy arrays weak-referenceable (the reference's WeakValueDictionary holds them directly, which
    numpy allows; plain lists / floats it does not -- those are simply not cached)."""
    __slots__ = ('value', '__weakref__')

    def __init__(self, value):
        self.value = value


class ExpBuffer:
    """exp_sender.py:10-59."""

    def __init__(self):
        self.exp_list = []
        self.ob_storage = {}

    def add(self, hash_dict, nonhash_dict):
        assert isinstance(hash_dict, dict) and isinstance(nonhash_dict, dict)
        exp = {}
        for key, values in hash_dict.items():
            assert not key.endswith('_hash'), 'do not manually append `_hash`'
            exp[key + '_hash'] = self._hash_nested(values)
        exp.update(nonhash_dict)
        self.exp_list.append(exp)

    def flush(self):
        binary = serialize((self.exp_list, self.ob_storage))
        self.exp_list, self.ob_storage = [], {}
        return binary

    def _hash_nested(self, values):
        if isinstance(values, list):
            return [self._hash_nested(v) for v in values]
        if isinstance(values, tuple):
            return tuple(self._hash_nested(v) for v in values)
        if isinstance(values, dict):
            return {k: self._hash_nested(v) for k, v in values.items()}
        if values is None:
            return None
        hsh = pyobj_hash(values)
        if hsh not in self.ob_storage:
            self.ob_storage[hsh] = values
        return hsh


class ExpSender:
    """exp_sender.py:62-98: buffer ``flush_iteration`` sends, then push one frame."""

    def __init__(self, *, host, port, flush_iteration):
        import zmq
        assert isinstance(flush_iteration, int) and flush_iteration >= 1
        self._ctx = zmq.Context.instance()
        self._sock = self._ctx.socket(zmq.PUSH)
        self._sock.connect('tcp://%s:%d' % (host, int(port)))
        self._exp_buffer = ExpBuffer()
        self._flush_iteration, self._count = flush_iteration, 0

    def send(self, hash_dict, nonhash_dict):
        self._exp_buffer.add(hash_dict=hash_dict, nonhash_dict=nonhash_dict)
        self._count += 1
        if self._count % self._flush_iteration == 0:
            binary = self._exp_buffer.flush()
            self._sock.send(binary)
            return binary_hash(binary)
        return None

    def close(self):
        self._sock.close(linger=1000)


def retrieve_storage(exp, storage, cache):
    """exp_collector.py:44-65: keys ending in ``_hash`` are replaced by the stored objects (suffix dropped)."""
    if isinstance(exp, list):
        return [retrieve_storage(e, storage, cache) for e in exp]
    if isinstance(exp, tuple):
        return tuple(retrieve_storage(e, storage, cache) for e in exp)
    if isinstance(exp, dict):
        out = {}
        for key, v in exp.items():
            if key.endswith('_hash'):
                out[key[:-len('_hash')]] = retrieve_storage(v, storage, cache)
            else:
                out[key] = retrieve_storage(v, storage, cache)
        return out
    if isinstance(exp, str):
        leaf = cache.get(exp)
        if leaf is None:
            if exp not in storage:
                return exp                                         # an ordinary string value, not a reference
            leaf = _Leaf(storage[exp])
            cache[exp] = leaf
        return leaf.value
    return exp


def inflate(binary, cache=None):
    """One frame -> list of experience dicts, exactly what the reference's collector hands to ``Replay.insert``."""
    exp_list, storage = deserialize(binary)
    cache = {} if cache is None else cache
    # only `_hash` keys are references; other string values must stay strings: resolve per key
    out = []
    for exp in exp_list:
        e = {}
        for key, v in exp.items():
            if key.endswith('_hash'):
                e[key[:-len('_hash')]] = retrieve_storage(v, storage, cache)
            else:
                e[key] = v
        out.append(e)
    return out


class ExperienceCollectorServer(threading.Thread):
    """exp_collector.py:7-65, same constructor: accepts frames from ExpSender-compatible actors on a PULL socket and calls
    ``exp_handler(exp)`` (normally ``replay.insert``) for every experience, in arrival order."""

    def __init__(self, host, port, exp_handler, load_balanced=False):
        super().__init__(daemon=True)
        self.host, self.port, self.load_balanced = host, int(port), load_balanced
        self._exp_handler = exp_handler
        self._cache = weakref.WeakValueDictionary()
        self._stop = threading.Event()
        self.frames = self.experiences = 0
        self.error = None

    def run(self):
        import zmq
        sock = zmq.Context.instance().socket(zmq.PULL)
        addr = 'tcp://%s:%d' % (self.host, self.port)
        (sock.connect if self.load_balanced else sock.bind)(addr)
        poller = zmq.Poller()
        poller.register(sock, zmq.POLLIN)
        try:
            while not self._stop.is_set():
                if not dict(poller.poll(50)):
                    continue
                for exp in inflate(sock.recv(), self._cache):
                    self._exp_handler(exp)
                    self.experiences += 1
                self.frames += 1
        except Exception as e:                                     # noqa: BLE001
            self.error = e
            raise
        finally:
            sock.close(linger=0)

    def stop(self):
        self._stop.set()
