"""oracle/replay.py -- TEST INFRASTRUCTURE: CPU restatement of the replay buffers and of the CPython
`random` calls they make.

Restates recovery_rl/replay_memory.py:11-33 (ReplayMemory) and :36-75 (ConstraintReplayMemory) of the
reference, plus the stdlib algorithm underneath (CPython 3.6-3.12 Lib/random.py `seed`, `sample`,
`_randbelow_with_getrandbits`; Modules/_randommodule.c `init_by_array`, `genrand_uint32`) -- a
third-party dependency of the reference that is not under /root/reference.

PINNED: against tests/golden/replay_idx.npz (index streams recorded from the reference's own
classes, plus known-answer vectors of CPython `random`), and -- because CPython itself is present on
every box -- against the live stdlib in tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
import numpy as np

N, M = 624, 397
MASK32 = 0xFFFFFFFF


def seed_key(seed):
    """Lib/random.py seed(int a): key = little-endian 32-bit limbs of abs(a), at least one limb."""
    a = abs(int(seed))
    limbs = []
    while True:
        limbs.append(a & MASK32)
        a >>= 32
        if a == 0:
            break
    return limbs


class MT19937(object):
    """Modules/_randommodule.c: init_genrand, init_by_array, genrand_uint32."""

    def __init__(self, seed):
        key = seed_key(seed)
        mt = [0] * N
        mt[0] = 19650218
        for i in range(1, N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & MASK32
        i, j = 1, 0
        for _ in range(max(N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & MASK32
            i += 1
            j += 1
            if i >= N:
                mt[0] = mt[N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & MASK32
            i += 1
            if i >= N:
                mt[0] = mt[N - 1]
                i = 1
        mt[0] = 0x80000000
        self.mt = mt
        self.index = N

    def state625(self):
        """uint32[625]: state words + index (the layout rrl_mt19937_seed_host produces)."""
        return np.array(self.mt + [self.index], np.uint32)

    def genrand_uint32(self):
        mt = self.mt
        if self.index >= N:
            for kk in range(N - M):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            for kk in range(N - M, N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk + 1] & 0x7FFFFFFF)
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            y = (mt[N - 1] & 0x80000000) | (mt[0] & 0x7FFFFFFF)
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.index = 0
        y = mt[self.index]
        self.index += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & MASK32

    def randbelow(self, n):
        """Lib/random.py _randbelow_with_getrandbits for 0 < n < 2**32."""
        k = int(n).bit_length()
        r = self.genrand_uint32() >> (32 - k)
        while r >= n:
            r = self.genrand_uint32() >> (32 - k)
        return r

    def sample_indices(self, n, k):
        """Lib/random.py sample(population, k) -> the chosen INDICES into the population."""
        if not 0 <= k <= n:
            raise ValueError("Sample larger than population or is negative")
        result = [0] * k
        setsize = 21
        if k > 5:
            p = 1
            while p < 3 * k:          # 4 ** ceil(log(3k, 4)); 3k is never a power of 4
                p *= 4
            setsize += p
        if n <= setsize:
            pool = list(range(n))
            for i in range(k):
                j = self.randbelow(n - i)
                result[i] = pool[j]
                pool[j] = pool[n - i - 1]
        else:
            selected = set()
            for i in range(k):
                j = self.randbelow(n)
                while j in selected:
                    j = self.randbelow(n)
                selected.add(j)
                result[i] = j
        return result


class SharedStream(object):
    """The ONE module-level generator both memories draw from; each constructor reseeds it
    (replay_memory.py:16,41)."""

    def __init__(self):
        self.rng = None

    def seed(self, seed):
        self.rng = MT19937(seed)


class ReplayMemory(object):
    """replay_memory.py:11-33.  Records are kept as fp32 rows [s0,s1,a0,a1,r,s2_0,s2_1,mask], the
    precision the reference converts to before any arithmetic (sac.py:185-190)."""

    def __init__(self, capacity, seed, stream):
        stream.seed(seed)
        self.stream = stream
        self.capacity = capacity
        self.buf = np.zeros((capacity, 8), np.float32)
        self.length = 0
        self.position = 0

    def push(self, state, action, reward, next_state, mask):
        self.buf[self.position] = (state[0], state[1], action[0], action[1], reward, next_state[0],
                                   next_state[1], mask)
        self._after_push()

    def _after_push(self):
        self.length = min(self.length + 1, self.capacity)
        self.position = (self.position + 1) % self.capacity

    def sample_slots(self, batch_size):
        return np.array(self.stream.rng.sample_indices(self.length, batch_size), np.int64)

    def gather(self, idx):
        b = self.buf[idx]
        return b[:, 0:2], b[:, 2:4], b[:, 4], b[:, 5:7], b[:, 7]

    def sample(self, batch_size):
        return self.gather(self.sample_slots(batch_size))

    def __len__(self):
        return self.length


class ConstraintReplayMemory(ReplayMemory):
    """replay_memory.py:36-75."""

    def __init__(self, capacity, seed, stream):
        ReplayMemory.__init__(self, capacity, seed, stream)
        self.pos_idx = np.zeros(capacity)

    def push(self, state, action, reward, next_state, mask):
        self.pos_idx[self.position] = reward
        ReplayMemory.push(self, state, action, reward, next_state, mask)

    def sample_slots(self, batch_size, pos_fraction=None):
        if pos_fraction is None:
            return ReplayMemory.sample_slots(self, batch_size)
        pos_size = int(batch_size * pos_fraction)
        neg_size = batch_size - pos_size
        pos_slots = np.argwhere(self.pos_idx).ravel()                       # whole capacity array (:58)
        neg_slots = np.argwhere((1 - self.pos_idx)[:self.length]).ravel()   # [:len(buffer)] (:62-65)
        pi = self.stream.rng.sample_indices(len(pos_slots), pos_size)
        ni = self.stream.rng.sample_indices(len(neg_slots), neg_size)
        return np.concatenate([pos_slots[pi], neg_slots[ni]]).astype(np.int64)

    def sample(self, batch_size, pos_fraction=None):
        return self.gather(self.sample_slots(batch_size, pos_fraction))
