"""Out-of-bounds writes: every caller-owned output of the fused steps sits between guard bands that must come back
untouched (compute-sanitizer is not available on the GPU pool, so the kernels' bounds are checked this way), over
ragged shapes (odd H*W, list counts that do not fill a warp or a tile) and every kernel family."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 1024  # elements on each side


def _guarded(shape, dtype, dev, sentinel):
    n = int(np.prod(shape))
    big = torch.full((n + 2 * GUARD,), sentinel, dtype=dtype, device=dev)
    return big, big[GUARD:GUARD + n].view(*shape)


def _intact(big, n, sentinel):
    return bool((big[:GUARD] == sentinel).all()) and bool((big[GUARD + n:] == sentinel).all())


CASES = [
    # strategy, K, R, B, H, W, Hm, Wm, holes, emit
    ("purely", 5, 1000, 3, 37, 29, 37, 29, False, True),
    ("purely", 5, 333, 3, 37, 29, 37, 29, True, True),
    ("purely", 5, 333, 3, 37, 29, 37, 29, True, False),      # valid-index accumulation
    ("purely", 1, 77, 2, 16, 16, 16, 16, False, True),
    ("purely", 16, 129, 2, 40, 24, 20, 12, True, True),      # down-scaled mask
    ("purely", 17, 65, 2, 40, 24, 40, 24, False, True),      # group-per-list kernels
    ("purely", 50, 131, 2, 40, 24, 40, 24, True, True),
    ("purely", 50, 131, 2, 40, 24, 40, 24, True, False),
    ("purely", 130, 33, 2, 40, 24, 40, 24, False, True),
    ("purely", 512, 9, 1, 40, 24, 40, 24, False, True),
    ("thresholded", 5, 333, 3, 37, 29, 37, 29, True, True),  # shared-memory selection
    ("thresholded", 5, 333, 3, 37, 29, 37, 29, True, False),
    ("information", 7, 2500, 2, 37, 29, 37, 29, False, True),   # 12500 candidates: radix selection + sort
    ("information", 7, 2500, 2, 37, 29, 37, 29, False, False),  # exact unordered selection
    ("masked", 9, 6000, 1, 37, 29, 37, 29, True, True),
    ("thresholded", 20, 200, 2, 37, 29, 37, 29, False, True),   # staged calls (K > 16)
]


@pytest.mark.parametrize("strategy,K,R,B,H,W,Hm,Wm,holes,emit", CASES)
def test_outputs_stay_inside_their_buffers(cuda_device, strategy, K, R, B, H, W, Hm, Wm, holes, emit):
    from pldepth_b200 import ops
    from pldepth_b200.step import FusedPLStep
    dev = cuda_device
    rs = np.random.RandomState(K * 1000 + R)
    gt = torch.from_numpy(rs.rand(B, H, W).astype(np.float32)).to(dev)
    mask_h = np.ones((B, Hm, Wm), np.float32)
    if holes:
        mask_h = (rs.rand(B, Hm, Wm) > 0.3).astype(np.float32)
    mask = torch.from_numpy(mask_h).to(dev)
    pred = torch.from_numpy(rs.randn(B, H, W, 1).astype(np.float32)).to(dev)
    S32, S64, SI = 12345.0, -54321.0, -77
    bigs = {}
    out = dict(key=None)
    for name, shape, dtype, sent in (("valid_flat", (B, Hm * Wm), torch.int32, SI), ("n_valid", (B,), torch.int32, SI),
                                     ("rankings", (B, R, K, 2), torch.float32, S32),
                                     ("grad", (B, H, W, 1), torch.float32, S32), ("loss", (1,), torch.float32, S32),
                                     ("loss_sum", (1,), torch.float64, S64)):
        big, view = _guarded(shape, dtype, dev, sent)
        bigs[name] = (big, int(np.prod(shape)), sent)
        out[name] = view
    if not emit:
        out["rankings"] = None
    step = FusedPLStep(K, R, seed=3, emit_rankings=emit, strategy=strategy)
    for _ in range(2):
        res = step.run(gt, mask, pred, out=out)
    torch.cuda.synchronize(dev)
    ops.check_status(dev)
    for name, (big, n, sent) in bigs.items():
        assert _intact(big, n, sent), "guard band of %s was written" % name
    assert np.isfinite(res["loss"].item())
    g = res["grad"]
    assert bool(torch.isfinite(g).all())
    if emit:
        r = res["rankings"]
        idx = r[..., 0]
        assert bool((idx >= 0).all()) and bool((idx < H * W).all())
        assert bool((r[..., :-1, 1] >= r[..., 1:, 1]).all())       # every list depth-descending
