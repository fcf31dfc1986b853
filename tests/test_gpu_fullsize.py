"""GPU parity at the FULL sizes of the BASELINE.json configs, through the headline entry point
(``FusedPLStep.run`` -> ``pld_fused_step``: lookup tables + fused list kernel, the path ``bench.py`` times).

For every named config the emitted rankings are checked as a whole (depth-descending, indices inside the valid
mask, label == gt at the index) and then handed to the oracle (``oracle.listmle_oracle.hourglass_nll``, fp64): the
loss and the ENTIRE dense gradient map must agree within 1e-5 (max norm relative to the largest gradient entry --
the tolerance ``north_star`` states for fp32).  The same step without materialised rankings (what a fused training
step runs; valid-index accumulation on holed masks) must reproduce the same loss bit for bit and the same dense
gradient within the same tolerance.

Reference semantics: sample_masked_rankings (pldepth/data/sampling.py:131-145) feeding
HourglassNegativeLogLikelihood (pldepth/losses/nll_loss.py:32-62).
"""
import numpy as np
import pytest
import torch

from oracle import listmle_oracle as lo

pytestmark = pytest.mark.gpu
RTOL = 1e-5

CASES = [
    # name, B, H, W, K, R, hole fraction, config id (synthetic-input seeds)
    ("C1", 4, 448, 448, 5, 1000, 0.0, 1),
    ("C2", 32, 448, 448, 5, 100000, 0.0, 2),
    ("C2-hole10", 32, 448, 448, 5, 100000, 0.1, 2),
    ("C3", 16, 448, 448, 50, 50000, 0.0, 3),
    ("C3-hole10", 16, 448, 448, 50, 50000, 0.1, 3),
    ("C4-loss-path", 64, 448, 448, 5, 1000, 0.0, 4),
    ("C5-slice", 4, 1024, 768, 10, 1000000, 0.0, 5),
]


def _inputs(B, H, W, hole, cfg):
    from pldepth_b200 import synth
    base = [synth.depth_map(H, W, 1000 * cfg + i) for i in range(min(B, 3))]
    gt = np.stack([np.roll(base[b % len(base)], 41 * b, axis=1) for b in range(B)])
    mask = np.stack([synth.valid_mask(H, W, 3000 * cfg + b, hole) for b in range(B)])
    pred = np.random.RandomState(2000 * cfg).standard_normal((B, H, W, 1)).astype(np.float32)
    return gt, mask, pred


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-300)


@pytest.mark.parametrize("name,B,H,W,K,R,hole,cfg", CASES, ids=[c[0] for c in CASES])
def test_headline_step_dense_parity_at_full_size(cuda_device, name, B, H, W, K, R, hole, cfg):
    from pldepth_b200.step import FusedPLStep
    gt, mask, pred = _inputs(B, H, W, hole, cfg)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))

    step = FusedPLStep(K, R, seed=cfg, emit_rankings=True)
    out = step.run(gt_d, mask_d, pred_d)
    step.check(cuda_device)
    rank = out["rankings"]
    assert tuple(rank.shape) == (B, R, K, 2) and tuple(out["grad"].shape) == (B, H, W, 1)

    # --- the rankings, all of them ---------------------------------------------------------------------------
    d = rank[..., 1]
    assert bool((d[:, :, :-1] >= d[:, :, 1:]).all()), "lists must be depth-descending"
    idx = rank[..., 0].long()
    assert int(idx.min()) >= 0 and int(idx.max()) < H * W
    flat_gt = gt_d.reshape(B, -1)
    assert bool((torch.gather(flat_gt, 1, idx.reshape(B, -1)).reshape(B, R, K) == d).all()), "label == gt[index]"
    assert bool((torch.gather(mask_d.reshape(B, -1), 1, idx.reshape(B, -1)) > 0).all()), "masked-out pixel drawn"
    nv = out["n_valid"].cpu().numpy()
    assert np.array_equal(np.abs(nv), (mask.reshape(B, -1) > 0).sum(axis=1))
    assert bool((nv < 0).all()) == (hole == 0.0)            # identity tables exactly for full masks

    # --- loss + the whole dense gradient against the oracle ----------------------------------------------------
    want_loss, want_grad, want_pl = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
    assert abs(out["loss"].item() - want_loss) <= RTOL * abs(want_loss), (out["loss"].item(), want_loss)
    assert abs(out["loss_sum"].item() - want_pl.sum()) <= RTOL * abs(want_pl.sum())
    got_grad = out["grad"].cpu().numpy()
    err = _rel(got_grad, want_grad)
    assert err <= RTOL, "%s dense gradient: relative error %.3e" % (name, err)
    # untouched pixels are exactly zero (grad is overwritten, never accumulated into)
    assert np.array_equal(got_grad == 0, want_grad == 0) or _rel(got_grad[want_grad == 0], 0 * got_grad[want_grad == 0] ) == 0

    # --- same step, rankings not materialised ------------------------------------------------------------------
    step2 = FusedPLStep(K, R, seed=cfg, emit_rankings=False)
    out2 = step2.run(gt_d, mask_d, pred_d)
    step2.check(cuda_device)
    assert out2["rankings"] is None
    assert out2["loss"].item() == out["loss"].item() and out2["loss_sum"].item() == out["loss_sum"].item()
    err2 = _rel(out2["grad"].cpu().numpy(), want_grad)
    assert err2 <= RTOL, "%s dense gradient without rankings: relative error %.3e" % (name, err2)


@pytest.mark.parametrize("strategy", ["thresholded", "information"])
def test_scored_step_dense_parity_at_config2_size(cuda_device, strategy):
    """The reference's default strategy (InformationScore, pldepth/PLDepth.py:44,100-101) and the validation
    sampler (Thresholded, hourglass_provider.py:22) at config-2 shapes: kept rankings -> oracle loss + dense gradient."""
    from pldepth_b200.step import FusedPLStep
    B, H, W, K, R = 32, 448, 448, 5, 100000
    gt, mask, pred = _inputs(B, H, W, 0.0, 2)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    step = FusedPLStep(K, R, seed=7, emit_rankings=True, strategy=strategy)
    out = step.run(gt_d, mask_d, pred_d)
    step.check(cuda_device)
    rank = out["rankings"]
    d = rank[..., 1]
    assert bool((d[:, :, :-1] >= d[:, :, 1:]).all())
    want_loss, want_grad, _ = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
    assert abs(out["loss"].item() - want_loss) <= RTOL * abs(want_loss)
    assert _rel(out["grad"].cpu().numpy(), want_grad) <= RTOL
    step2 = FusedPLStep(K, R, seed=7, emit_rankings=False, strategy=strategy)
    out2 = step2.run(gt_d, mask_d, pred_d)
    step2.check(cuda_device)
    # the kept SET is the same, its order is not observable without rankings: loss equal up to summation order
    assert abs(out2["loss"].item() - want_loss) <= RTOL * abs(want_loss)
    assert _rel(out2["grad"].cpu().numpy(), want_grad) <= RTOL
