"""GPU: evaluation list generators (generic_ranking_provider.py:80-111, 180-215) against the goldens written by the
unmodified reference and against the oracle on larger seeded inputs, bit for bit including the RNG state afterwards."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import eval_lists_oracle as eo
from oracle.reference_loader import DictModelParams

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAIR_FILES = sorted(glob.glob(os.path.join(GOLD, "eval_pairs_*.npz")))
RANK_FILES = sorted(glob.glob(os.path.join(GOLD, "eval_rankings_*.npz")))


def dataset(gts, channels=3):
    return [(np.zeros(gts[i].shape + (channels,), np.float32), gts[i][..., None]) for i in range(len(gts))]


def same_state(a, b):
    return np.array_equal(a[1], b[1]) and a[2] == b[2]


@pytest.mark.parametrize("rng", ["numpy", "mt19937"])
@pytest.mark.parametrize("path", PAIR_FILES, ids=[os.path.basename(p)[5:-4] for p in PAIR_FILES])
def test_ordinal_pairs_reproduce_reference_goldens(cuda_device, path, rng):
    from pldepth_b200.eval_lists import GenericHourglassPairRelationDataProvider
    g = np.load(path)
    thr = None if float(g["threshold"]) < 0 else float(g["threshold"])
    seed = int(g["seed"])
    mp = DictModelParams(val_rankings_per_img=int(g["n_pairs"]), dataset="synthetic")
    prov = GenericHourglassPairRelationDataProvider(mp, seed, bool(g["invert"]), threshold=thr, rng=rng)
    np.random.seed(seed)
    got = prov.generate_ordinal_pairs(dataset(g["gts"]), invert_relation_sign=bool(g["invert"]))
    assert got.dtype == np.float32 and np.array_equal(got, g["pairs"])
    if rng == "numpy":        # the global stream stands where the reference left it
        rs = np.random.RandomState(seed)
        rs.randint(0, 2 ** 32, size=int(g["consumed"]), dtype=np.uint32)
        assert same_state(np.random.get_state(), rs.get_state())


@pytest.mark.parametrize("rng", ["numpy", "mt19937"])
@pytest.mark.parametrize("path", RANK_FILES, ids=[os.path.basename(p)[5:-4] for p in RANK_FILES])
def test_rankings_reproduce_reference_goldens(cuda_device, path, rng):
    from pldepth_b200.eval_lists import GenericHourglassRankingDataProvider
    g = np.load(path)
    seed = int(g["seed"])
    prov = GenericHourglassRankingDataProvider(DictModelParams(dataset="synthetic"), int(g["K"]), seed, bool(g["invert"]),
                                               rng=rng)
    np.random.seed(seed)
    got = prov.generate_rankings(dataset(g["gts"]), invert_relation_sign=bool(g["invert"]),
                                 val_rankings_per_img=int(g["n_lists"]))
    assert got.dtype == np.float32 and np.array_equal(got, g["rankings"])
    if rng == "numpy":
        rs = np.random.RandomState(seed)
        rs.randint(0, 2 ** 32, size=int(g["consumed"]), dtype=np.uint32)
        assert same_state(np.random.get_state(), rs.get_state())


def test_larger_seeded_cases_equal_the_oracle_and_chain_the_stream(cuda_device):
    """Two generators called back to back on one stream (as a validation + test provider pair would), mixed image
    sizes in one dataset, both promotions of the relation arithmetic."""
    from pldepth_b200.eval_lists import GenericHourglassPairRelationDataProvider, GenericHourglassRankingDataProvider
    rs = np.random.RandomState(5)

    def maps(n, H, W):
        return np.stack([(0.05 * np.power(1.011, rs.permutation(H * W) % 200) * (1 + 1e-6 * rs.permutation(H * W)))
                         .astype(np.float32).reshape(H, W) for _ in range(n)])
    a, b = maps(3, 56, 72), maps(2, 31, 45)
    ds = dataset(a) + dataset(b, channels=1)[:0] + [(np.zeros((31, 45), np.float32), b[i]) for i in range(2)]
    for promotion in ("nep50", "legacy"):
        mp = DictModelParams(val_rankings_per_img=700, dataset="x")
        pp = GenericHourglassPairRelationDataProvider(mp, 9, True, threshold=0.03, promotion=promotion)
        rp = GenericHourglassRankingDataProvider(mp, 9, 9, False)
        np.random.seed(9)
        pairs = pp.generate_ordinal_pairs(ds, invert_relation_sign=True)
        lists = rp.generate_rankings(ds, val_rankings_per_img=150)
        end = np.random.get_state()
        ref = np.random.RandomState(9)
        if promotion == "nep50":
            want_pairs = np.concatenate([eo.generate_ordinal_pairs(a, 700, 0.03, True, rng=ref),
                                         eo.generate_ordinal_pairs(b, 700, 0.03, True, rng=ref)])
        else:   # NumPy 1.x value-based casting: float32 scalar + python float -> float64 ratio
            want_pairs = np.concatenate([eo.generate_ordinal_pairs(a.astype(np.float64), 700, 0.03, True, rng=ref),
                                         eo.generate_ordinal_pairs(b.astype(np.float64), 700, 0.03, True, rng=ref)])
        want_lists = np.concatenate([eo.generate_rankings(a, 9, 150, False, rng=ref).reshape(-1, 9, 2),
                                     eo.generate_rankings(b, 9, 150, False, rng=ref).reshape(-1, 9, 2)])
        assert np.array_equal(pairs, want_pairs)
        assert np.array_equal(lists.reshape(-1, 9, 2), want_lists)
        assert same_state(end, ref.get_state())
        assert (pairs[:, :, 2] == 0).sum() > 10          # the threshold band is exercised


def test_provide_val_dataset_seeds_like_the_reference(cuda_device, tmp_path):
    from pldepth_b200.eval_lists import GenericHourglassRankingDataProvider
    rs = np.random.RandomState(1)
    gts = np.stack([((rs.permutation(400) + 0.5) / 400).astype(np.float32).reshape(20, 20) for _ in range(2)])
    cfg = {"DATA": {"CACHE_PATH_PREFIX": str(tmp_path)}}
    os.makedirs(str(tmp_path / "ranking_cache"))
    prov = GenericHourglassRankingDataProvider(DictModelParams(dataset="syn"), 4, 21, True, save_rankings_on_disk=True,
                                               config=cfg)
    _, lists = prov.provide_val_dataset(dataset(gts))
    want = eo.generate_rankings(gts, 4, 100, True, rng=np.random.RandomState(21))
    assert np.array_equal(lists, want)
    assert os.path.isfile(str(tmp_path / "ranking_cache" / "syn_val_100_21_4.npy"))
    np.random.seed(0)                                     # second call comes from the cache, stream untouched
    st = np.random.get_state()
    _, again = prov.provide_val_dataset(dataset(gts))
    assert np.array_equal(again, want)
    with pytest.raises(NotImplementedError):
        prov.provide_train_dataset(None)
    with pytest.raises(ValueError):
        GenericHourglassRankingDataProvider(DictModelParams(dataset="syn"), 4, 21, True, rng="philox")
