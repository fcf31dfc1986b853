"""Synthetic inputs of the BASELINE.json shapes: tie-free rank-transformed
ground-truth depth, valid masks, N(0,1) predictions.  NumPy only; deterministic per seed."""
import numpy as np

CONFIGS = {
    "C1": dict(B=4, H=448, W=448, K=5, R=1000),
    "C2": dict(B=32, H=448, W=448, K=5, R=100000),
    "C3": dict(B=16, H=448, W=448, K=50, R=50000),
    "C4": dict(B=64, H=448, W=448, K=5, R=1000),
    "C5": dict(B=256, H=1024, W=768, K=10, R=1000000),
}


def depth_map(H, W, seed):
    """Smooth random field (a few low-frequency cosines + noise), rank-transformed to
    (rank + 0.5) / HW as float32: values in (0, 1), all distinct while HW < 2^23."""
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64) / H, np.arange(W, dtype=np.float64) / W, indexing="ij")
    f = np.zeros((H, W))
    for _ in range(4):
        fy, fx = rng.uniform(0.5, 3.0, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=2)
        f += rng.uniform(0.5, 1.0) * np.cos(2 * np.pi * fy * yy + ph[0]) * np.cos(2 * np.pi * fx * xx + ph[1])
    f += 0.05 * rng.standard_normal((H, W))
    order = np.argsort(f.reshape(-1), kind="stable")
    ranks = np.empty(H * W, dtype=np.int64)
    ranks[order] = np.arange(H * W)
    return ((ranks + 0.5) / (H * W)).astype(np.float32).reshape(H, W)


def valid_mask(Hm, Wm, seed, hole_fraction=0.0):
    m = np.ones((Hm, Wm), dtype=np.float32)
    if hole_fraction > 0:
        rng = np.random.RandomState(seed)
        area = hole_fraction * Hm * Wm
        h = max(1, int(round(np.sqrt(area * Hm / Wm))))
        w = max(1, int(round(area / h)))
        h, w = min(h, Hm), min(w, Wm)
        r0 = rng.randint(0, Hm - h + 1)
        c0 = rng.randint(0, Wm - w + 1)
        m[r0:r0 + h, c0:c0 + w] = 0
    return m


def prediction(H, W, seed):
    return np.random.RandomState(seed).standard_normal((H, W, 1)).astype(np.float32)


def batch(config_id, B, H, W, hole_fraction=0.0, first_image=0):
    gt = np.stack([depth_map(H, W, 1000 * config_id + first_image + b) for b in range(B)])
    mask = np.stack([valid_mask(H, W, 3000 * config_id + first_image + b, hole_fraction) for b in range(B)])
    pred = np.stack([prediction(H, W, 2000 * config_id + first_image + b) for b in range(B)])
    return gt, mask, pred
