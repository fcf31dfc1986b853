"""Generate the committed golden vectors by running the UNMODIFIED reference sampler.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Each case stores the inputs (gt, mask, seed, K, R, strategy, factor) and what the reference's
own ``sample_masked_point_batch`` / ``sample_masked_rankings`` returned under
``np.random.seed(seed)``, plus the number of MT19937 words the call consumed (measured by
advancing a copy of the start state until it equals the end state).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def tie_free_gt(H, W, seed):
    rng = np.random.RandomState(seed)
    return ((rng.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W)


def near_threshold_gt(H, W, seed):
    """Distinct depths on a geometric ladder with ratio ~1.0101 so that many adjacent pairs of a
    sorted list fall on either side of the 1.03 "equal" threshold (depth_utils.py:16-21)."""
    rng = np.random.RandomState(seed)
    ladder = (0.05 * np.power(1.0101, np.arange(H * W) % 300) * (1 + 1e-6 * (np.arange(H * W) // 300)))
    return rng.permutation(ladder).astype(np.float32).reshape(H, W)


def words_consumed(st0, st1, limit):
    rs = np.random.RandomState()
    rs.set_state(st0)
    def same(a, b):
        return np.array_equal(a[1], b[1]) and a[2] == b[2]
    if same(rs.get_state(), st1):
        return 0
    # advance word by word in blocks
    n = 0
    while n < limit:
        rs.randint(0, 2 ** 32, size=1, dtype=np.uint32)
        n += 1
        if same(rs.get_state(), st1):
            return n
    raise RuntimeError("could not match end state")


CASES = [
    # name, strategy class, H, W, Hm, Wm, K, R, factor, seed, gt kind, hole
    ("purely_k5", "PurelyMaskedRandomSamplingStrategy", 24, 32, 24, 32, 5, 40, None, 11, "perm", True),
    ("purely_core_k5_f1", "PurelyMaskedRandomSamplingStrategy", 24, 32, 24, 32, 5, 40, 1.0, 12, "perm", True),
    ("purely_scaled_k3", "PurelyMaskedRandomSamplingStrategy", 24, 32, 12, 8, 3, 30, 1.0, 13, "perm", True),
    ("purely_k20", "PurelyMaskedRandomSamplingStrategy", 20, 20, 20, 20, 20, 12, 1.0, 14, "perm", False),
    ("masked_k5", "MaskedRandomSamplingStrategy", 24, 32, 24, 32, 5, 40, None, 15, "perm", True),
    ("thresholded_k5", "ThresholdedMaskedRandomSamplingStrategy", 24, 32, 24, 32, 5, 40, None, 16, "ladder", True),
    ("thresholded_k9", "ThresholdedMaskedRandomSamplingStrategy", 24, 32, 24, 32, 9, 30, None, 17, "ladder", False),
    ("information_k5", "InformationScoreBasedSampling", 24, 32, 24, 32, 5, 30, None, 18, "ladder", True),
    ("information_k12", "InformationScoreBasedSampling", 24, 32, 24, 32, 12, 20, None, 19, "perm", False),
    ("information_k3_full", "InformationScoreBasedSampling", 16, 16, 16, 16, 3, 25, None, 20, "perm", False),
    # long lists (group-per-list kernels, NumPy pairwise summation beyond 8 / 128 terms)
    ("purely_k50", "PurelyMaskedRandomSamplingStrategy", 40, 48, 40, 48, 50, 20, 1.0, 21, "perm", True),
    ("thresholded_k33", "ThresholdedMaskedRandomSamplingStrategy", 40, 48, 40, 48, 33, 16, None, 22, "perm", True),
    ("information_k130", "InformationScoreBasedSampling", 40, 48, 40, 48, 130, 6, None, 23, "perm", False),
    # score-based strategies on down-scaled / holed masks, the shortest list, a larger call
    ("thresholded_scaled_k4", "ThresholdedMaskedRandomSamplingStrategy", 24, 32, 12, 16, 4, 30, None, 24, "ladder", True),
    ("information_holes_k8", "InformationScoreBasedSampling", 32, 24, 32, 24, 8, 25, None, 25, "perm", True),
    ("masked_k2", "MaskedRandomSamplingStrategy", 40, 48, 40, 48, 2, 16, None, 26, "ladder", True),
    ("information_scaled_k6", "InformationScoreBasedSampling", 24, 32, 8, 8, 6, 20, None, 27, "ladder", False),
    ("thresholded_k16_r60", "ThresholdedMaskedRandomSamplingStrategy", 48, 64, 48, 64, 16, 60, None, 29, "ladder", True),
]


def main():
    S = rl.load_reference_sampling()
    for name, cls, H, W, Hm, Wm, K, R, factor, seed, kind, hole in CASES:
        gt = tie_free_gt(H, W, seed) if kind == "perm" else near_threshold_gt(H, W, seed)
        mask = np.ones((Hm, Wm), np.float32)
        if hole:
            mask[Hm // 4: Hm // 2, Wm // 8: Wm // 2] = 0
            mask[0, 0] = 0
        image = np.zeros((H, W, 3), np.float32)
        mp = rl.DictModelParams(ranking_size=K)
        strat = getattr(S, cls)(mp)
        np.random.seed(seed)
        st0 = np.random.get_state()
        if factor is None:
            out = strat.sample_masked_point_batch(image, mask, gt, R)
            used_factor = -1.0
        else:
            out = strat.sample_masked_point_batch(image, mask, gt, R, factor) if cls != "PurelyMaskedRandomSamplingStrategy" \
                else strat.sample_masked_rankings(image, mask, gt, R, factor)[0]
            used_factor = float(factor)
        st1 = np.random.get_state()
        consumed = words_consumed(st0, st1, 200000)
        if cls != "PurelyMaskedRandomSamplingStrategy":
            # a golden is only a pin if the reference's (unstable) argsort had no tied scores to break
            from oracle import sampler_oracle as so
            strat_name = {"MaskedRandomSamplingStrategy": "masked", "ThresholdedMaskedRandomSamplingStrategy": "thresholded",
                          "InformationScoreBasedSampling": "information"}[cls]
            _, _, scores = so.sample_masked_point_batch(strat_name, (H, W), mask, gt, R, K, factor,
                                                        rng=np.random.RandomState(seed))
            if np.unique(scores).size != scores.size:
                print("  note: %s has tied candidate scores (sum of adjacent differences telescopes to max - min); "
                      "the reference's unstable argsort may order them either way" % name)
        np.savez_compressed(os.path.join(OUT, "sampler_%s.npz" % name), gt=gt, mask=mask, seed=seed, K=K, R=R,
                            factor=used_factor, strategy=cls, rankings=np.asarray(out, dtype=np.float32),
                            consumed=consumed, numpy_version=np.__version__)
        print(name, out.shape, "consumed", consumed)


if __name__ == "__main__":
    main()
