"""MT19937 restated from the published algorithm (Matsumoto & Nishimura 1998) -- TEST
INFRASTRUCTURE ONLY.  Pinned against ``numpy.random.RandomState`` (same generator the
reference draws from at sampling.py:113): ``tests/test_oracle_rng.py`` checks
``init_genrand`` against ``RandomState(seed).get_state()`` and the tempered output against
``randint(0, 2**32, dtype=uint32)``.  Used to check the CUDA MT19937 stream generator.
"""
import numpy as np

N, M_ = 624, 397
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


def init_genrand(seed):
    """Knuth-style LCG state fill used by ``np.random.seed(int)`` (numpy ``mt19937_seed``)."""
    mt = np.zeros(N, dtype=np.uint32)
    x = int(seed) & 0xFFFFFFFF
    mt[0] = x
    for i in range(1, N):
        x = (1812433253 * (x ^ (x >> 30)) + i) & 0xFFFFFFFF
        mt[i] = x
    return mt


def twist(mt):
    """One full state regeneration (624 words), sequential form."""
    mt = [int(v) for v in mt]
    for i in range(N):
        y = (mt[i] & UPPER) | (mt[(i + 1) % N] & LOWER)
        mt[i] = mt[(i + M_) % N] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
    return np.array(mt, dtype=np.uint32)


def temper(y):
    y = np.asarray(y, dtype=np.uint32).copy()
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


def raw_words(state, pos, n):
    """n tempered outputs starting from (state[624], pos in [0, 624]); returns (words, state, pos)."""
    out = np.empty(n, dtype=np.uint32)
    state = np.asarray(state, dtype=np.uint32).copy()
    k = 0
    while k < n:
        if pos >= N:
            state = twist(state)
            pos = 0
        take = min(N - pos, n - k)
        out[k:k + take] = temper(state[pos:pos + take])
        pos += take
        k += take
    return out, state, pos
