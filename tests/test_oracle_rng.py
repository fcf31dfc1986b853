"""CPU: the RNG models the parity tests rely on, pinned against NumPy and published KATs."""
import numpy as np

from oracle import mt19937_oracle as mto
from oracle import sampler_oracle as so
from tests import philox_model as pm


def test_init_genrand_matches_numpy_seed():
    for seed in (0, 1, 3, 12345, 2 ** 32 - 1):
        st = np.random.RandomState(seed).get_state()
        assert np.array_equal(st[1], mto.init_genrand(seed))
        assert st[2] == 624


def test_mt19937_stream_matches_numpy():
    for seed in (0, 7):
        want = np.random.RandomState(seed).randint(0, 2 ** 32, size=2000, dtype=np.uint32)
        got, state, pos = mto.raw_words(mto.init_genrand(seed), 624, 2000)
        assert np.array_equal(got, want)
        more = np.random.RandomState(seed)
        more.randint(0, 2 ** 32, size=2000, dtype=np.uint32)
        want2 = more.randint(0, 2 ** 32, size=100, dtype=np.uint32)
        got2, _, _ = mto.raw_words(state, pos, 100)
        assert np.array_equal(got2, want2)


def test_masked_rejection_is_numpy_randint():
    for M in (1, 2, 3, 5, 1000, 3846, 4096, 4097, 200704, 786432):
        rs = np.random.RandomState(5)
        st0 = rs.get_state()
        want = rs.randint(M, size=777)
        st1 = rs.get_state()
        rs.set_state(st0)
        raw = so.raw_words_from_state(rs, 4000)
        got, consumed = so.masked_rejection(raw, M, 777)
        assert np.array_equal(got, want)
        rs.set_state(st0)
        so.raw_words_from_state(rs, consumed)
        st2 = rs.get_state()
        assert np.array_equal(st1[1], st2[1]) and st1[2] == st2[2]


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    r = pm.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = pm.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = pm.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_selection_uniform_and_in_range():
    sel = pm.draw_selection(seed=9, offset=2, image=1, n=4000, K=5, M=1000)
    assert sel.min() >= 0 and sel.max() < 1000
    hist = np.bincount(sel.reshape(-1), minlength=1000)
    chi2 = ((hist - 20.0) ** 2 / 20.0).sum()
    assert 800 < chi2 < 1200          # 999 dof, +-4.5 sigma
