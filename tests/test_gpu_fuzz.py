"""GPU: randomized differential test -- random shapes, ranking sizes, masks and strategies; the one-call step,
the staged calls and the oracle must agree (rankings bit-exact, loss / gradient within 1e-5)."""
import os

import numpy as np
import pytest
import torch

from oracle import listmle_oracle as lo
from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu


def random_case(rs):
    B = int(rs.randint(1, 5))
    H, W = int(rs.randint(5, 72)), int(rs.randint(5, 72))
    scaled = rs.rand() < 0.25
    # masks coarser AND finer than the image (x_scale > 1 / < 1, sampling.py:124-129)
    Hm, Wm = (int(rs.randint(3, 2 * H)), int(rs.randint(3, 2 * W))) if scaled else (H, W)
    K = int(rs.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 17, 20, 31, 33, 48, 64, 65, 100]))
    n = int(rs.randint(1, 700 if K <= 16 else 60))
    gt = np.stack([((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W) for _ in range(B)])
    kind = rs.randint(0, 3)
    mask = np.ones((B, Hm, Wm), np.float32)
    if kind >= 1:
        mask = (rs.rand(B, Hm, Wm) > rs.uniform(0.05, 0.7)).astype(np.float32)
        mask[:, 0, 0] = 1.0                  # never empty
        if kind == 2 and B > 1:
            mask[0] = 1.0                    # mixed batch: one full mask
    pred = (rs.randn(B, H, W, 1) * rs.uniform(0.2, 3.0)).astype(np.float32)
    return B, H, W, Hm, Wm, K, n, gt, mask, pred


N_CORE = int(os.environ.get("PLD_FUZZ_CASES", "40"))
N_SCORED = max(6, (N_CORE * 3) // 5)


@pytest.mark.parametrize("seed", range(N_CORE))
def test_random_core_step(cuda_device, seed):
    from pldepth_b200 import ops
    rs = np.random.RandomState(1000 + seed)
    B, H, W, Hm, Wm, K, n, gt, mask, pred = random_case(rs)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    s, off, base = int(rs.randint(0, 2 ** 31)), int(rs.randint(0, 1000)), int(rs.randint(0, 50))
    loss, loss_sum, grad, rank, pl, nv = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=s, offset=off, image_base=base,
                                                        want_per_list=True)
    vf, nv2 = ops.mask_compact(mask_d, H, W)
    rank2, sel = ops.sample_lists_philox(gt_d, vf, nv2, K, n, s, off, base, want_sel=True)
    ops.check_status(cuda_device)
    assert torch.equal(nv, nv2) and torch.equal(rank, rank2)
    rank_h, sel_h = rank.cpu().numpy(), sel.cpu().numpy()
    for b in range(B):
        want = so.rankings_from_selection(sel_h[b].reshape(-1), so.valid_flat_indices(mask[b], (H, W)), gt[b], K)
        assert np.array_equal(rank_h[b], want)
    want_loss, want_grad, want_pl = lo.hourglass_nll(rank_h, pred, B, K)
    assert abs(loss.item() - want_loss) <= 1e-5 * max(abs(want_loss), 1e-12)
    if K > 1:
        err = np.abs(grad.cpu().numpy() - want_grad).max() / np.abs(want_grad).max()
        assert err <= 1e-5
        assert np.abs(pl.cpu().numpy() - want_pl).max() <= 1e-5 * np.abs(want_pl).max()
    # no materialised rankings (valid-index accumulation on holed masks)
    loss4, _, grad4, _, _, _ = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=s, offset=off, image_base=base,
                                              want_rankings=False)
    assert loss4.item() == loss.item()
    if K > 1:
        assert np.abs(grad4.cpu().numpy() - want_grad).max() / np.abs(want_grad).max() <= 1e-5
    # the fed-ranking loss (Keras signature path) on the same lists
    loss3, _, grad3, _ = ops.listmle_fwd_bwd(rank, pred_d, B, K, 1.0 / (B * n))
    assert abs(loss3.item() - want_loss) <= 1e-5 * max(abs(want_loss), 1e-12)


@pytest.mark.parametrize("seed", range(N_SCORED))
def test_random_scored_step(cuda_device, seed):
    from pldepth_b200 import ops
    rs = np.random.RandomState(5000 + seed)
    B, H, W, Hm, Wm, K, n, gt, mask, pred = random_case(rs)
    K = int(rs.choice([2, 3, 5, 7, 8, 9, 11, 16]))
    # every fourth case has more than 8192 candidates per image (radix selection + segmented sort instead of the
    # shared-memory sort)
    n = int(rs.randint(8193, 12000)) if seed % 4 == 3 else int(rs.randint(2, 500))
    R = int(rs.randint(1, n + 1))
    strategy = ["masked", "thresholded", "information"][seed % 3]
    promotion = ["nep50", "legacy"][(seed // 3) % 2]
    if rs.rand() < 0.3:      # quantised depths: heavy score ties
        gt = (np.floor(gt * 6) / 6 + 0.05).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    out = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, promotion, seed=seed, offset=3,
                                want_order=True)
    vf, nv = ops.mask_compact(mask_d, H, W)
    cand, sel = ops.sample_lists_philox(gt_d, vf, nv, K, n, seed, 3, 0, want_sel=True)
    mm = ops.gt_minmax(gt_d) if strategy == "information" else None
    scores = ops.score_lists(cand, strategy, 0.03, -1000, promotion, mm)
    top, order = ops.select_top(scores, cand, R, want_order=True)
    ops.check_status(cuda_device)
    assert torch.equal(out["order"], order) and torch.equal(out["rankings"], top)
    # and against the oracle, image by image
    sel_h, top_h = sel.cpu().numpy(), top.cpu().numpy()
    for b in range(B):
        cand_b = so.rankings_from_selection(sel_h[b].reshape(-1), so.valid_flat_indices(mask[b], (H, W)), gt[b], K)
        if strategy == "information":
            sc = so.score_information(cand_b, gt[b], 0.03, -1000, promotion)
        else:
            sc = so.score_adjacent_differences(cand_b, 0.03 if strategy == "thresholded" else None, -1000, promotion)
        want_top, _ = so.select_top(cand_b, sc, R)
        assert np.array_equal(top_h[b], want_top)
    want_loss, want_grad, _ = lo.hourglass_nll(top_h, pred, B, K)
    assert abs(out["loss"].item() - want_loss) <= 1e-5 * max(abs(want_loss), 1e-12)
    err = np.abs(out["grad"].cpu().numpy() - want_grad).max() / np.abs(want_grad).max()
    assert err <= 1e-5
    # unordered selection (rankings not materialised): same kept set
    out3 = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, promotion, seed=seed, offset=3,
                                 want_rankings=False, want_order=True)
    assert torch.equal(torch.sort(out3["order"], dim=1).values, torch.sort(order, dim=1).values)
    assert abs(out3["loss"].item() - want_loss) <= 1e-5 * max(abs(want_loss), 1e-12)
    assert np.abs(out3["grad"].cpu().numpy() - want_grad).max() / np.abs(want_grad).max() <= 1e-5
