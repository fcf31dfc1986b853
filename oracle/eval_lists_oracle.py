"""NumPy restatement of the reference's evaluation list generators -- TEST INFRASTRUCTURE ONLY.

Follows pldepth/data/providers/generic_ranking_provider.py:80-111 (``generate_ordinal_pairs``) and 180-215
(``generate_rankings``).  Pinned against the UNMODIFIED reference file run behind the TensorFlow stub of
``oracle/reference_loader.py``: ``tests/golden/eval_*.npz`` (written by ``tests/golden/make_golden_eval.py``).
``rng`` is a legacy ``numpy.random.RandomState`` (or the ``numpy.random`` module for the global state the
reference itself uses)."""
import numpy as np


def depth_relation(z0, z1, threshold):
    """pldepth/data/depth_utils.py:5-21 on NumPy scalars (float32 in, so NumPy >= 2 computes the ratio in float32)."""
    if threshold is None:
        return 1 if z0 > z1 else (-1 if z0 < z1 else 0)
    eps = 1e-10
    ratio = (z0 + eps) / (z1 + eps)
    if ratio >= 1 + threshold:
        return 1
    if ratio <= 1 / (1 + threshold):
        return -1
    return 0


def generate_ordinal_pairs(gts, n_pairs, threshold=0.03, invert_relation_sign=False, rng=np.random):
    """gts: [N,H,W] float32 -> float32 [N, n_pairs, 5] = (point0, point1, relation, z0, z1);
    four ``randint`` calls per pair with bounds H, W, H, W (generic_ranking_provider.py:91-95)."""
    gts = np.asarray(gts)
    N, H, W = gts.shape
    out = np.zeros([N, n_pairs, 5], np.float32)
    for i in range(N):
        gt = gts[i]
        for j in range(n_pairs):
            x0 = rng.randint(H)
            y0 = rng.randint(W)
            x1 = rng.randint(H)
            y1 = rng.randint(W)
            z0, z1 = gt[x0, y0], gt[x1, y1]
            rel = depth_relation(z0, z1, threshold)
            if invert_relation_sign:
                rel *= -1
            out[i, j] = np.array([x0 * W + y0, x1 * W + y1, rel, z0, z1])
    return out


def generate_rankings(gts, ranking_size, n_lists=100, invert_relation_sign=False, rng=np.random):
    """gts: [N,H,W] float32 -> float32 [N, n_lists, K, 2] = (flat index, depth); K ``randint(0, H*W)`` draws per list
    (generic_ranking_provider.py:188-196); ordered by depth descending, or -- inverted -- by original depth ascending
    with the depth stored as 1 / (depth + 1) (float64 arithmetic, one rounding to float32; 198-209).
    Ties are broken by a stable rule (reversed stable argsort resp. stable argsort); the reference's default
    ``np.argsort`` is unstable, so parity is claimed on tie-free depth maps."""
    gts = np.asarray(gts)
    N = gts.shape[0]
    K = int(ranking_size)
    out = np.zeros([N, n_lists, K, 2], np.float32)
    for i in range(N):
        gt = gts[i].reshape([-1])
        for r in range(n_lists):
            pts = np.zeros([K, 2])
            for k in range(K):
                s = rng.randint(0, len(gt))
                pts[k, 0] = s
                pts[k, 1] = gt[s]
            if invert_relation_sign:
                order = np.argsort(pts[:, 1], kind="stable")
                pts[:, 1] = 1 / (pts[:, 1] + 1)
            else:
                order = np.argsort(pts[:, 1], kind="stable")[::-1]
            out[i, r] = pts[order]
    return out
