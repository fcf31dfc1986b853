"""Generate golden vectors for the evaluation list generators by running the UNMODIFIED reference file
pldepth/data/providers/generic_ranking_provider.py (generate_ordinal_pairs 80-111, generate_rankings 180-215).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_eval.py
Each case stores the inputs (gts, seed, sizes, flags), the array the reference returned under
``np.random.seed(seed)`` and the number of MT19937 words it consumed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402
from tests.golden.make_golden import near_threshold_gt, tie_free_gt, words_consumed  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

PAIR_CASES = [
    # name, N, H, W, pairs per image, threshold (None = plain comparison), invert, seed, gt kind
    ("pairs_thr", 3, 24, 40, 60, 0.03, False, 41, "ladder"),
    ("pairs_thr_inverted", 2, 33, 17, 50, 0.03, True, 42, "ladder"),
    ("pairs_plain", 2, 16, 64, 40, None, False, 43, "perm"),        # power-of-two bounds: no rejected words
    ("pairs_wide_thr", 2, 30, 30, 50, 0.25, True, 44, "ladder"),
]
RANK_CASES = [
    # name, N, H, W, K, lists per image, invert, seed
    ("rankings_k5", 3, 24, 40, 5, 30, False, 51),
    ("rankings_k5_inverted", 2, 24, 40, 5, 30, True, 52),
    ("rankings_k12_inverted", 2, 32, 32, 12, 20, True, 53),          # H*W a power of two
    ("rankings_k40", 2, 40, 48, 40, 10, False, 54),
    ("rankings_k1", 2, 10, 12, 1, 25, True, 55),
]


def gts_of(kind, N, H, W, seed):
    f = tie_free_gt if kind == "perm" else near_threshold_gt
    return np.stack([f(H, W, seed * 10 + i) for i in range(N)])


def main():
    ref = rl.load_reference_eval_providers()
    for name, N, H, W, n_pairs, thr, invert, seed, kind in PAIR_CASES:
        gts = gts_of(kind, N, H, W, seed)
        ds = rl.ListDataset([(np.zeros((H, W, 1), np.float32), gts[i][..., None]) for i in range(N)])
        mp = rl.DictModelParams(val_rankings_per_img=n_pairs, dataset="synthetic")
        prov = ref.GenericHourglassPairRelationDataProvider(mp, seed, invert, threshold=thr)
        np.random.seed(seed)
        st0 = np.random.get_state()
        out = prov.generate_ordinal_pairs(ds, invert_relation_sign=invert)
        consumed = words_consumed(st0, np.random.get_state(), 200000)
        np.savez_compressed(os.path.join(OUT, "eval_%s.npz" % name), gts=gts, seed=seed, n_pairs=n_pairs,
                            threshold=-1.0 if thr is None else thr, invert=invert, pairs=out, consumed=consumed,
                            numpy_version=np.__version__)
        print(name, out.shape, "consumed", consumed, "relations", np.unique(out[:, :, 2], return_counts=True))
    for name, N, H, W, K, n_lists, invert, seed in RANK_CASES:
        gts = gts_of("perm", N, H, W, seed)
        ds = rl.ListDataset([(np.zeros((H, W, 1), np.float32), gts[i][..., None]) for i in range(N)])
        mp = rl.DictModelParams(dataset="synthetic")
        prov = ref.GenericHourglassRankingDataProvider(mp, K, seed, invert)
        np.random.seed(seed)
        st0 = np.random.get_state()
        out = prov.generate_rankings(ds, invert_relation_sign=invert, val_rankings_per_img=n_lists)
        consumed = words_consumed(st0, np.random.get_state(), 200000)
        np.savez_compressed(os.path.join(OUT, "eval_%s.npz" % name), gts=gts, seed=seed, K=K, n_lists=n_lists,
                            invert=invert, rankings=out, consumed=consumed, numpy_version=np.__version__)
        print(name, out.shape, "consumed", consumed)


if __name__ == "__main__":
    main()
