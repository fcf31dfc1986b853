"""CPU: the evaluation-list oracle against the goldens written by the unmodified reference
(tests/golden/make_golden_eval.py) and, where /root/reference exists, against the live reference file."""
import glob
import os

import numpy as np
import pytest

from oracle import eval_lists_oracle as eo
from oracle import reference_loader as rl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAIR_FILES = sorted(glob.glob(os.path.join(GOLD, "eval_pairs_*.npz")))
RANK_FILES = sorted(glob.glob(os.path.join(GOLD, "eval_rankings_*.npz")))


def words_used(seed, state_after, limit=200000):
    rs = np.random.RandomState(int(seed))
    for n in range(limit):
        st = rs.get_state()
        if np.array_equal(st[1], state_after[1]) and st[2] == state_after[2]:
            return n
        rs.randint(0, 2 ** 32, size=1, dtype=np.uint32)
    raise AssertionError("end state not reached")


def test_goldens_exist():
    assert len(PAIR_FILES) >= 4 and len(RANK_FILES) >= 5


@pytest.mark.parametrize("path", PAIR_FILES, ids=[os.path.basename(p)[5:-4] for p in PAIR_FILES])
def test_ordinal_pairs_oracle_reproduces_reference(path):
    g = np.load(path)
    thr = None if float(g["threshold"]) < 0 else float(g["threshold"])
    rs = np.random.RandomState(int(g["seed"]))
    got = eo.generate_ordinal_pairs(g["gts"], int(g["n_pairs"]), thr, bool(g["invert"]), rng=rs)
    assert got.dtype == np.float32 and np.array_equal(got, g["pairs"])
    assert words_used(g["seed"], rs.get_state()) == int(g["consumed"])


@pytest.mark.parametrize("path", RANK_FILES, ids=[os.path.basename(p)[5:-4] for p in RANK_FILES])
def test_rankings_oracle_reproduces_reference(path):
    g = np.load(path)
    rs = np.random.RandomState(int(g["seed"]))
    got = eo.generate_rankings(g["gts"], int(g["K"]), int(g["n_lists"]), bool(g["invert"]), rng=rs)
    assert got.dtype == np.float32 and np.array_equal(got, g["rankings"])
    assert words_used(g["seed"], rs.get_state()) == int(g["consumed"])
    d = got[..., 1]
    if got.shape[2] > 1:       # inverted lists: original depth ascending = stored 1/(d+1) descending, too
        assert (np.diff(d, axis=2) <= 0).all()


@pytest.mark.skipif(not rl.reference_available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("invert", [False, True])
def test_oracle_equals_live_reference(invert):
    ref = rl.load_reference_eval_providers()
    rs = np.random.RandomState(7)
    N, H, W = 2, 18, 22
    gts = np.stack([((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W) for _ in range(N)])
    ds = rl.ListDataset([(np.zeros((H, W, 3), np.float32), gts[i][..., None]) for i in range(N)])
    prov = ref.GenericHourglassPairRelationDataProvider(rl.DictModelParams(val_rankings_per_img=25, dataset="x"), 3,
                                                        invert, threshold=0.1)
    np.random.seed(3)
    want = prov.generate_ordinal_pairs(ds, invert_relation_sign=invert)
    end = np.random.get_state()
    rs2 = np.random.RandomState(3)
    got = eo.generate_ordinal_pairs(gts, 25, 0.1, invert, rng=rs2)
    assert np.array_equal(got, want) and np.array_equal(rs2.get_state()[1], end[1])
    prov = ref.GenericHourglassRankingDataProvider(rl.DictModelParams(dataset="x"), 7, 4, invert)
    np.random.seed(4)
    want = prov.generate_rankings(ds, invert_relation_sign=invert, val_rankings_per_img=12)
    got = eo.generate_rankings(gts, 7, 12, invert, rng=np.random.RandomState(4))
    assert np.array_equal(got, want)
