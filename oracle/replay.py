"""Replay semantics (reference: surreal/replay/fifo_replay.py:6-48, uniform_replay.py:6-74) and the
CPython ``random.randint`` index stream they depend on."""
import random
from collections import deque


class FIFO:
    """fifo_replay.py: deque(maxlen=memory_size+3) -> silently drops the OLDEST on overflow (:27);
    sample pops from the left in arrival order (:37-39); ready when len >= batch_size (:44-45)."""

    def __init__(self, memory_size, batch_size):
        self.memory_size = memory_size
        self.batch_size = batch_size
        self.q = deque(maxlen=memory_size + 3)

    def insert(self, x):
        self.q.append(x)

    def sample(self, batch_size):
        assert batch_size <= self.memory_size
        return [self.q.popleft() for _ in range(batch_size)]

    def ready(self):
        return len(self.q) >= self.batch_size

    def __len__(self):
        return len(self.q)


class Uniform:
    """uniform_replay.py: k-th insert lands in slot k % memory_size (:36-41); sample draws
    ``batch`` i.i.d. ``randint(0, len-1)`` WITH replacement (:43-47); ready when len > start (:70-71)."""

    def __init__(self, memory_size, sampling_start_size, rng=None):
        self.memory_size = memory_size
        self.start = sampling_start_size
        self.mem = []
        self.next_idx = 0
        self.rng = rng if rng is not None else random      # module-level Random, like the reference

    def insert(self, x):
        if self.next_idx >= len(self.mem):
            self.mem.append(x)
        else:
            self.mem[self.next_idx] = x
        self.next_idx = (self.next_idx + 1) % self.memory_size

    def sample_indices(self, batch_size):
        return [self.rng.randint(0, len(self.mem) - 1) for _ in range(batch_size)]

    def sample(self, batch_size):
        return [self.mem[i] for i in self.sample_indices(batch_size)]

    def ready(self):
        return len(self.mem) > self.start

    def __len__(self):
        return len(self.mem)


# ---- independent model of CPython's generator (validates the C++ MT19937 in the product) --------
class MT19937:
    """Textbook MT19937 with CPython's ``init_by_array`` seeding (Modules/_randommodule.c) --
    written from the published algorithm; checked against ``random.Random`` in the tests."""
    N, M = 624, 397

    def __init__(self, seed):
        key = []
        s = abs(int(seed))
        while True:
            key.append(s & 0xFFFFFFFF)
            s >>= 32
            if s == 0:
                break
        self.mt = [0] * self.N
        self._init_genrand(19650218)
        i, j = 1, 0
        mt = self.mt
        for _ in range(max(self.N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & 0xFFFFFFFF
            i += 1
            j += 1
            if i >= self.N:
                mt[0] = mt[self.N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(self.N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & 0xFFFFFFFF
            i += 1
            if i >= self.N:
                mt[0] = mt[self.N - 1]
                i = 1
        mt[0] = 0x80000000
        self.idx = self.N

    def _init_genrand(self, s):
        self.mt[0] = s & 0xFFFFFFFF
        for i in range(1, self.N):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF

    def genrand_uint32(self):
        mt, N, M = self.mt, self.N, self.M
        if self.idx >= N:
            for k in range(N):
                y = (mt[k] & 0x80000000) | (mt[(k + 1) % N] & 0x7FFFFFFF)
                mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def randbelow(self, m):
        """random.Random._randbelow_with_getrandbits for m < 2**32: k = m.bit_length();
        draw getrandbits(k) = genrand_uint32() >> (32-k) until < m."""
        k = int(m).bit_length()
        while True:
            r = self.genrand_uint32() >> (32 - k)
            if r < m:
                return r

    def randint(self, a, b):
        return a + self.randbelow(b - a + 1)
