"""GPU parity tests, stages 2-3 (gather + ListMLE NLL + gradient scatter-add) and the fused
step, through the C ABI.  Tolerance: 1e-5 relative (BASELINE.json north_star), measured on the
loss and on the gradient map in the max norm relative to the largest gradient entry."""
import numpy as np
import pytest
import torch

from oracle import listmle_oracle as lo
from oracle import sampler_oracle as so

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def make_problem(B, H, W, K, R, seed, sorted_lists=True, dup=True):
    rs = np.random.RandomState(seed)
    pred = (rs.randn(B, H, W, 1) * 1.5).astype(np.float32)
    idx = rs.randint(0, H * W, size=(B, R, K))
    if dup and K >= 2:
        idx[:, 0, 1] = idx[:, 0, 0]
    depth = rs.permutation(B * R * K).reshape(B, R, K).astype(np.float64)
    depth = (depth + 0.5) / (B * R * K)
    if sorted_lists:
        depth = np.sort(depth, axis=2)[:, :, ::-1]
    y_true = np.stack([idx.astype(np.float32), depth.astype(np.float32)], axis=-1)
    return y_true, pred


def assert_close(got, want, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    denom = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / denom
    assert err <= RTOL, "%s: relative error %.3e > %.1e" % (what, err, RTOL)


@pytest.mark.parametrize("K", [1, 2, 3, 5, 8, 10, 16, 17, 32, 50, 64, 100, 128, 250, 256, 500, 512])
@pytest.mark.parametrize("sorted_lists", [True, False])
def test_loss_and_gradient_match_oracle(cuda_device, K, sorted_lists):
    from pldepth_b200 import ops
    B, H, W = 3, 24, 20
    R = 130 if K <= 64 else 19
    y_true, pred = make_problem(B, H, W, K, R, 10 * K + sorted_lists, sorted_lists)
    want_loss, want_grad, want_pl = lo.hourglass_nll(y_true, pred, B, K)
    loss, loss_sum, grad, per_list = ops.listmle_fwd_bwd(torch.from_numpy(y_true).to(cuda_device),
                                                         torch.from_numpy(pred).to(cuda_device), B, K,
                                                         1.0 / (B * R), want_grad=True, want_per_list=True)
    ops.check_status(cuda_device)
    assert_close(loss.item(), want_loss, "loss")
    assert_close(loss_sum.item(), want_pl.sum(), "loss_sum")
    assert_close(per_list.cpu().numpy(), want_pl, "per-list NLL")
    assert grad.shape == pred.shape
    assert_close(grad.cpu().numpy(), want_grad, "gradient")


def test_known_answer_through_keras_signature(cuda_device):
    from pldepth_b200.losses import HourglassNegativeLogLikelihood
    # two lists on a 1x3 map: scores (0, ln3, ln2); orders (1,2,0) and (2,0,1)
    pred = torch.tensor([[[[0.0], [np.log(3.0)], [np.log(2.0)]]]], dtype=torch.float32, device=cuda_device,
                        requires_grad=True)
    y_true = torch.tensor([[[[1, 0.9], [2, 0.5], [0, 0.1]], [[2, 0.9], [0, 0.5], [1, 0.1]]]], dtype=torch.float32,
                          device=cuda_device)
    loss_fn = HourglassNegativeLogLikelihood(ranking_size=3, batch_size=1)
    val = loss_fn(y_true, pred)
    # list 1: P = 3/6 * 2/3 -> nll = ln2 + ln1.5 ; list 2: P = 2/6 * 1/4 -> nll = ln3 + ln4
    want = 0.5 * ((np.log(2) + np.log(1.5)) + (np.log(3) + np.log(4)))
    assert abs(val.item() - want) < 1e-6
    val.backward()
    assert abs(pred.grad.sum().item()) < 1e-6


def test_invalid_labels_and_ties(cuda_device):
    from pldepth_b200 import ops
    B, H, W, K, R = 2, 10, 10, 6, 50
    y_true, pred = make_problem(B, H, W, K, R, 5, sorted_lists=False)
    rs = np.random.RandomState(0)
    neg = rs.rand(B, R, K) < 0.2
    y_true[..., 1][neg] = -1.0
    y_true[:, 1, 2, 1] = y_true[:, 1, 4, 1] = 0.5        # a tie between valid labels (stable rule)
    y_true[:, 3, :, 1] = -2.0                            # a list with no valid label at all
    want_loss, want_grad, want_pl = lo.hourglass_nll(y_true, pred, B, K)
    loss, _, grad, per_list = ops.listmle_fwd_bwd(torch.from_numpy(y_true).to(cuda_device),
                                                  torch.from_numpy(pred).to(cuda_device), B, K, 1.0 / (B * R),
                                                  want_per_list=True)
    assert_close(per_list.cpu().numpy(), want_pl, "per-list NLL")
    assert_close(loss.item(), want_loss, "loss")
    assert_close(grad.cpu().numpy(), want_grad, "gradient")


def test_bad_index_raises(cuda_device):
    from pldepth_b200 import ops
    y_true, pred = make_problem(1, 4, 4, 3, 5, 1)
    y_true[0, 2, 1, 0] = 16.0
    ops.listmle_fwd_bwd(torch.from_numpy(y_true).to(cuda_device), torch.from_numpy(pred).to(cuda_device), 1, 3, 1.0)
    with pytest.raises(IndexError):
        ops.check_status(cuda_device)


def test_reductions_autograd_and_validation_R(cuda_device):
    from pldepth_b200.losses import HourglassNegativeLogLikelihood
    B, H, W, K = 2, 12, 12, 5
    for R in (7, 30):                                   # R is inferred per call (nll_loss.py:58)
        y_true, pred = make_problem(B, H, W, K, R, R)
        yt = torch.from_numpy(y_true).to(cuda_device)
        for red, factor in (("auto", 1.0), ("sum", float(B * R))):
            p = torch.from_numpy(pred).to(cuda_device).requires_grad_(True)
            fn = HourglassNegativeLogLikelihood(K, B, reduction=red)
            v = fn(yt, p)
            (v * 3.0).backward()
            want_loss, want_grad, want_pl = lo.hourglass_nll(y_true, pred, B, K, reduction=red)
            assert_close(v.item(), want_loss, "loss " + red)
            assert_close(p.grad.cpu().numpy(), 3.0 * want_grad, "grad " + red)
        none = HourglassNegativeLogLikelihood(K, B, reduction="none")(yt, torch.from_numpy(pred).to(cuda_device))
        assert none.shape == (B * R, 1)
        assert_close(none.cpu().numpy()[:, 0], want_pl, "per-list")
    with pytest.raises(NotImplementedError):
        HourglassNegativeLogLikelihood(K, B, lambda_weight=object())


@pytest.mark.parametrize("K,n", [(5, 1000), (10, 300), (50, 100), (200, 20)])
def test_fused_equals_sample_then_loss(cuda_device, K, n):
    """Fused step == Philox sampling followed by the loss on its rankings; both == oracle."""
    from pldepth_b200 import ops
    from tests.test_gpu_sampler import make_maps
    B, H, W = 3, 36, 44
    gt, mask = make_maps(H, W, H, W, K, B)
    pred = np.random.RandomState(K).randn(B, H, W, 1).astype(np.float32)
    gt_d, pred_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    loss, loss_sum, grad, rank, per_list = ops.fused_sample_loss_bwd(gt_d, vf, nv, pred_d, K, n, seed=77, offset=3,
                                                                     want_per_list=True)
    rank2, _ = ops.sample_lists_philox(gt_d, vf, nv, K, n, 77, 3, 0)
    assert torch.equal(rank, rank2)
    loss2, _, grad2, pl2 = ops.listmle_fwd_bwd(rank2, pred_d, B, K, 1.0 / (B * n), want_per_list=True)
    assert torch.equal(per_list, pl2)
    assert abs(loss.item() - loss2.item()) <= 1e-6 * abs(loss2.item())
    assert_close(grad.cpu().numpy(), grad2.cpu().numpy(), "fused vs two-step gradient")
    want_loss, want_grad, _ = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
    assert_close(loss.item(), want_loss, "loss")
    assert_close(grad.cpu().numpy(), want_grad, "gradient")
    # forward-only and no-rankings variants
    l3, _, g3, r3, _ = ops.fused_sample_loss_bwd(gt_d, vf, nv, pred_d, K, n, seed=77, offset=3, want_rankings=False,
                                                 want_grad=False)
    assert g3 is None and r3 is None and l3.item() == loss.item()


@pytest.mark.parametrize("K,n", [(1, 50), (2, 400), (5, 1500), (8, 700), (16, 300), (17, 90), (50, 150), (200, 30),
                                 (512, 9)])
@pytest.mark.parametrize("geometry", ["full", "holes", "scaled", "finer", "finer-rows"])
def test_one_call_step_equals_staged_calls(cuda_device, K, n, geometry):
    """pld_fused_step (lookup tables + fused kernel) == pld_mask_compact + pld_fused_sample_loss_bwd:
    rankings and per-list NLL bit for bit, gradient within tolerance, for full / holed / down-scaled masks and for
    masks FINER than the image (x_scale < 1, sampling.py:124-129: several valid mask pixels map to one image pixel, whose
    gradient contributions must accumulate -- with and without materialised rankings)."""
    from pldepth_b200 import ops
    from tests.test_gpu_sampler import make_maps
    B, H, W = 3, 40, 48
    Hm, Wm = {"scaled": (20, 16), "finer": (60, 96), "finer-rows": (80, 48)}.get(geometry, (H, W))
    gt, mask = make_maps(H, W, Hm, Wm, 7 * K, B, hole=(geometry != "full"))
    if geometry == "holes":
        mask[1] = 1.0                           # mix of full and holed images in one batch
    pred = np.random.RandomState(K).randn(B, H, W, 1).astype(np.float32)
    gt_d, pred_d, mask_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, pred, mask))
    loss, loss_sum, grad, rank, pl, nv = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=5, offset=9, image_base=2,
                                                        want_per_list=True)
    vf, nv2 = ops.mask_compact(mask_d, H, W)
    assert torch.equal(nv, nv2)
    loss2, ls2, grad2, rank2, pl2 = ops.fused_sample_loss_bwd(gt_d, vf, nv2, pred_d, K, n, seed=5, offset=9,
                                                              image_base=2, want_per_list=True)
    ops.check_status(cuda_device)
    assert torch.equal(rank, rank2)
    assert torch.equal(pl, pl2)
    assert loss.item() == loss2.item() and loss_sum.item() == ls2.item()
    assert_close(grad.cpu().numpy(), grad2.cpu().numpy(), "one-call vs staged gradient")
    want_loss, want_grad, _ = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
    assert_close(loss.item(), want_loss, "loss")
    assert_close(grad.cpu().numpy(), want_grad, "gradient")
    for b in range(B):                          # never a masked-out pixel
        valid = set(so.valid_flat_indices(mask[b], (H, W)).tolist())
        assert set(rank[b, :, :, 0].long().flatten().tolist()) <= valid
    # without materialising the rankings (valid-index gradient accumulation for holed masks): same loss, same gradient
    loss3, ls3, grad3, rank3, pl3, nv3 = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=5, offset=9, image_base=2,
                                                        want_rankings=False, want_per_list=True)
    assert rank3 is None and torch.equal(pl3, pl) and loss3.item() == loss.item() and torch.equal(nv3, nv)
    assert_close(grad3.cpu().numpy(), want_grad, "gradient without emitted rankings")


@pytest.mark.parametrize("K,n", [(17, 300), (24, 200), (33, 300), (50, 400), (56, 100), (64, 200)])
@pytest.mark.parametrize("geometry", ["full", "holes"])
@pytest.mark.parametrize("depths", ["clustered", "tied", "tiny"])
def test_long_lists_order_on_the_full_depth(cuda_device, K, n, geometry, depths):
    """ranking_size 17..64 (pld_lists_tab.cu) orders by a truncated depth prefix | draw slot and re-orders a list
    exactly when two of its depths agree in the prefix.  Depth maps built to collide: a few hundred float32 neighbours (runs of 64 share a prefix),
    a handful of exactly tied values (ties: later draw first, DESIGN.md "Oracle"), and zero / tiny depths.  The
    emitted rankings must equal the oracle's on the same draws and the staged kernel's (exact 64-bit keys)."""
    from pldepth_b200 import ops
    B, H, W = 2, 24, 28
    rs = np.random.RandomState(K + len(depths))
    if depths == "clustered":
        steps = rs.randint(0, 300, size=(B, H, W))
        gt = (np.float32(0.5) + steps.astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    elif depths == "tied":
        gt = rs.choice(np.array([0.125, 0.25, 0.2500001, 0.7, 0.70000005], np.float32), size=(B, H, W)).astype(np.float32)
    else:
        gt = (rs.randint(0, 80, size=(B, H, W)).astype(np.float32) * np.float32(2.0 ** -140)).astype(np.float32)
    mask = (rs.rand(B, H, W) > (0.3 if geometry == "holes" else -1)).astype(np.float32)
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    gt_d, pred_d, mask_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, pred, mask))
    loss, _, grad, rank, pl, nv = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=21, offset=4, want_per_list=True)
    vf, nv2 = ops.mask_compact(mask_d, H, W)
    _, sel = ops.sample_lists_philox(gt_d, vf, nv2, K, n, 21, 4, 0, want_sel=True, want_rankings=False)
    _, _, grad2, rank2, _ = ops.fused_sample_loss_bwd(gt_d, vf, nv2, pred_d, K, n, seed=21, offset=4)
    ops.check_status(cuda_device)
    assert torch.equal(rank, rank2)
    rank_h = rank.cpu().numpy()
    for b in range(B):
        want = so.rankings_from_selection(sel[b].cpu().numpy().reshape(-1), so.valid_flat_indices(mask[b], (H, W)),
                                          gt[b], K)
        assert np.array_equal(rank_h[b], want)
    want_loss, want_grad, want_pl = lo.hourglass_nll(rank_h, pred, B, K)
    assert_close(loss.item(), want_loss, "loss")
    assert_close(pl.cpu().numpy(), want_pl, "per-list NLL")
    assert_close(grad.cpu().numpy(), want_grad, "gradient")
    # rankings not materialised (valid-index layout on holed masks): same lists, same loss
    loss3, _, grad3, _, pl3, _ = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=21, offset=4, want_rankings=False,
                                                want_per_list=True)
    assert torch.equal(pl3, pl) and loss3.item() == loss.item()
    assert_close(grad3.cpu().numpy(), want_grad, "gradient without emitted rankings")


@pytest.mark.parametrize("H,W,K,n", [(40, 48, 5, 900), (33, 37, 5, 500), (17, 129, 20, 60), (64, 64, 50, 40)])
@pytest.mark.parametrize("emit", [True, False])
def test_uint8_and_bool_masks_equal_float_masks(cuda_device, H, W, K, n, emit):
    """pld_fused_step_m8: a uint8 / bool mask (nonzero = valid, the HR-WSI masks as stored) gives bit for bit the step
    of the float32 mask -- emitted and not (bit-mask mode: odd sizes whose rows are neither 16-pixel nor 16-byte
    aligned, a full-mask image and an empty-hole image in the same batch)."""
    from pldepth_b200.step import FusedPLStep
    B = 3
    rs = np.random.RandomState(H * W + K)
    gt = rs.rand(B, H, W).astype(np.float32)
    mask = (rs.rand(B, H, W) > 0.3).astype(np.float32)
    mask[1] = 1.0
    mask[2, :, : W // 2] = 0.0
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    gt_d, pred_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    outs = []
    for m in (torch.from_numpy(mask), torch.from_numpy((mask * 200).astype(np.uint8)), torch.from_numpy(mask > 0)):
        st = FusedPLStep(K, n, seed=13, emit_rankings=emit)
        out = st.run(gt_d, m.to(cuda_device), pred_d)
        st.check(cuda_device)
        outs.append(out)
    for o in outs[1:]:
        assert torch.equal(o["n_valid"], outs[0]["n_valid"])
        assert o["loss"].item() == outs[0]["loss"].item()
        if emit:
            assert torch.equal(o["rankings"], outs[0]["rankings"])
        assert_close(o["grad"].cpu().numpy(), outs[0]["grad"].cpu().numpy(), "gradient, uint8 vs float32 mask")
    # ... and the float-mask step is the oracle's
    if emit:
        want_loss, want_grad, _ = lo.hourglass_nll(outs[0]["rankings"].cpu().numpy(), pred, B, K)
        assert_close(outs[0]["loss"].item(), want_loss, "loss")
        assert_close(outs[0]["grad"].cpu().numpy().reshape(B, -1), want_grad.reshape(B, -1), "gradient")
    g = outs[0]["grad"].cpu().numpy().reshape(B, H, W)
    assert (g[mask == 0] == 0).all()          # masked-out pixels: exactly zero, written (not left over)


def test_loss_is_deterministic(cuda_device):
    from pldepth_b200 import ops
    y_true, pred = make_problem(4, 64, 64, 5, 5000, 3)
    yt, p = torch.from_numpy(y_true).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    vals = {ops.listmle_fwd_bwd(yt, p, 4, 5, 1.0 / 20000, want_grad=False)[0].item() for _ in range(5)}
    assert len(vals) == 1


def test_full_size_config2_properties(cuda_device):
    """BASELINE config 2 (B32, 448^2, K5, R100k) at full size: size-independent properties --
    lists depth-descending with in-range indices, per-list gradients sum to zero so the dense
    gradient sums to ~0, loss equals the mean of the per-list NLL, and a 2k-list sample of the
    emitted rankings reproduces the oracle's per-list NLL."""
    from pldepth_b200 import ops, synth
    B, H, W, K, R = 32, 448, 448, 5, 100000
    rs = np.random.RandomState(0)
    gt1 = synth.depth_map(H, W, 2000)
    gt = np.stack([np.roll(gt1, 37 * b, axis=1) for b in range(B)])
    mask = np.ones((B, H, W), np.float32)
    mask[:, 100:160, 50:300] = 0
    pred = rs.standard_normal((B, H, W, 1)).astype(np.float32)
    gt_d, pred_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    loss, loss_sum, grad, rank, per_list = ops.fused_sample_loss_bwd(gt_d, vf, nv, pred_d, K, R, seed=2, offset=0,
                                                                     want_per_list=True)
    ops.check_status(cuda_device)
    assert int(nv[0].item()) == H * W - 60 * 250
    d = rank[..., 1]
    assert bool((d[:, :, :-1] >= d[:, :, 1:]).all())
    idx = rank[..., 0].long()
    assert int(idx.min()) >= 0 and int(idx.max()) < H * W
    rows, cols = idx // W, idx % W
    assert not bool(((rows >= 100) & (rows < 160) & (cols >= 50) & (cols < 300)).any())   # masked pixels never drawn
    assert bool((torch.gather(gt_d.reshape(B, -1), 1, idx.reshape(B, -1)).reshape(B, R, K) == d).all())
    assert abs(loss.item() - per_list.double().mean().item()) <= 1e-6 * abs(loss.item())
    gsum = grad.double().sum().item()
    gabs = grad.double().abs().sum().item()
    assert abs(gsum) <= 1e-5 * gabs
    pick = rs.randint(0, R, size=2000)
    sub = rank[5, pick].cpu().numpy()[None]
    _, _, want_pl = lo.hourglass_nll(sub, pred[5:6], 1, K)
    assert_close(per_list.reshape(B, R)[5, pick].cpu().numpy(), want_pl, "per-list NLL at full size")


def test_host_pipelined_step_matches_device_step(cuda_device):
    """HostPipelinedStep (pinned host in, loss + gradient out, overlapped copies) returns, step by
    step, what FusedPLStep computes on device-resident inputs."""
    from pldepth_b200.step import FusedPLStep, HostPipelinedStep
    from tests.test_gpu_sampler import make_maps
    B, H, W, K, R = 3, 32, 40, 5, 700
    runner = HostPipelinedStep(K, R, B, H, W, seed=11, device=cuda_device)
    ref = FusedPLStep(K, R, seed=11)
    tickets, wants = [], []
    for i in range(5):
        gt, mask = make_maps(H, W, H, W, 50 + i, B)
        pred = np.random.RandomState(i).randn(B, H, W, 1).astype(np.float32)
        tickets.append(runner.submit(gt, mask, pred))
        out = ref.run(*(torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred)))
        wants.append((out["loss"].item(), out["grad"].cpu().numpy().copy()))
        if i >= 1:                                  # results stay available for `slots` steps
            loss, grad = runner.result(tickets[i - 1])
            assert loss == wants[i - 1][0]
            assert_close(grad.numpy(), wants[i - 1][1], "pipelined gradient")
    loss, grad = runner.result(tickets[-1])
    assert loss == wants[-1][0]
    with pytest.raises(ValueError):
        runner.result(tickets[0])


def test_deterministic_mode_is_bit_reproducible(cuda_device):
    """With pld_ctx_set_deterministic the dense gradient is identical run to run, identical between the
    one-call step and the staged calls (different kernels / launch geometry), and still within
    tolerance of the oracle."""
    from pldepth_b200 import ops
    from pldepth_b200._lib import Context
    from tests.test_gpu_sampler import make_maps
    B, H, W, K, n = 3, 48, 40, 5, 4000          # 60k points on 1920 pixels: heavy accumulation per pixel
    gt, mask = make_maps(H, W, H, W, 21, B)
    pred = np.random.RandomState(3).randn(B, H, W, 1).astype(np.float32)
    gt_d, pred_d, mask_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, pred, mask))
    ctx = Context.current(cuda_device.index or 0)
    ctx.set_deterministic(True)
    try:
        grads = []
        for _ in range(3):
            _, _, g, rank, _, _ = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=8)
            grads.append(g.clone())
        assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
        vf, nv = ops.mask_compact(mask_d, H, W)
        _, _, g2, rank2, _ = ops.fused_sample_loss_bwd(gt_d, vf, nv, pred_d, K, n, seed=8)
        assert torch.equal(rank, rank2) and torch.equal(grads[0], g2)
        _, _, g3, _ = ops.listmle_fwd_bwd(rank, pred_d, B, K, 1.0 / (B * n))
        assert torch.equal(grads[0], g3)
        _, want_grad, _ = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
        assert_close(grads[0].cpu().numpy(), want_grad, "deterministic gradient")
    finally:
        ctx.set_deterministic(False)


def test_fused_step_object_with_strategy(cuda_device):
    from pldepth_b200 import ops
    from pldepth_b200.step import FusedPLStep
    from tests.test_gpu_sampler import make_maps
    B, H, W, K, R = 2, 32, 32, 5, 200
    gt, mask = make_maps(H, W, H, W, 2, B)
    pred = np.random.RandomState(0).randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    st = FusedPLStep(K, R, seed=3, strategy="thresholded")
    out = st.run(gt_d, mask_d, pred_d)
    ref = ops.fused_step_scored(mask_d, gt_d, pred_d, K, int(R * 1.5), R, "thresholded", seed=3, offset=0)
    assert torch.equal(out["rankings"], ref["rankings"]) and out["loss"].item() == ref["loss"].item()
    # long lists: same one-call entry point (group-per-list kernels)
    st20 = FusedPLStep(20, 50, seed=3, strategy="masked")
    out20 = st20.run(gt_d, mask_d, pred_d)
    want_loss, want_grad, _ = lo.hourglass_nll(out20["rankings"].cpu().numpy(), pred, B, 20)
    assert_close(out20["loss"].item(), want_loss, "K=20 scored loss")
    assert_close(out20["grad"].cpu().numpy(), want_grad, "K=20 scored gradient")
    with pytest.raises(ValueError):
        FusedPLStep(5, R, strategy="masked", candidate_factor=0.5)


@pytest.mark.parametrize("B,H,W,K,n", [(1, 7, 9, 3, 1), (2, 5, 5, 1, 40), (1, 31, 33, 5, 255), (3, 17, 19, 16, 257),
                                       (2, 64, 64, 5, 0)])
def test_edge_shapes_one_call_step(cuda_device, B, H, W, K, n):
    """Ragged / tiny / empty shapes: odd pixel counts (scalar table path), one list, K = 1, n = 0."""
    from pldepth_b200 import ops
    from tests.test_gpu_sampler import make_maps
    gt, mask = make_maps(H, W, H, W, H + W, B, hole=False)
    pred = np.random.RandomState(0).randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    loss, loss_sum, grad, rank, pl, nv = ops.fused_step(mask_d, gt_d, pred_d, K, n, seed=1, scale=1.0 / max(1, B * n),
                                                        want_per_list=True)
    ops.check_status(cuda_device)
    assert (nv < 0).all()
    if n == 0:
        assert loss.item() == 0 and rank.numel() == 0 and float(grad.abs().sum()) == 0
        return
    want_loss, want_grad, want_pl = lo.hourglass_nll(rank.cpu().numpy(), pred, B, K)
    assert_close(pl.cpu().numpy(), want_pl, "per-list") if K > 1 else None
    assert abs(loss.item() - want_loss) <= 1e-5 * max(abs(want_loss), 1e-30)
    if K > 1:
        assert_close(grad.cpu().numpy(), want_grad, "gradient")
    else:
        assert float(grad.abs().max()) == 0 and loss.item() == 0     # K = 1: NLL and gradient are exactly 0


def test_scored_step_edges(cuda_device):
    """R == n (keep everything, pure ordering), R == 1, K == 1 (all scores equal: order = index descending)."""
    from pldepth_b200 import ops
    from tests.test_gpu_sampler import make_maps
    B, H, W = 2, 20, 20
    gt, mask = make_maps(H, W, H, W, 3, B)
    gt_d, mask_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    for K, n, R in ((5, 64, 64), (5, 300, 1), (1, 50, 20)):
        out = ops.fused_step_scored(mask_d, gt_d, None, K, n, R, "thresholded", seed=2, want_order=True)
        vf, nv = ops.mask_compact(mask_d, H, W)
        cand, _ = ops.sample_lists_philox(gt_d, vf, nv, K, n, 2, 0, 0)
        top, order = ops.select_top(ops.score_lists(cand, "thresholded"), cand, R, want_order=True)
        assert torch.equal(out["order"], order) and torch.equal(out["rankings"], top)
        if K == 1:
            assert order[0].tolist() == list(range(n - 1, n - 1 - R, -1))


def test_maximum_map_size(cuda_device):
    """PLD_MAX_PIXELS = 2^23 pixels (flat indices are packed into 23 bits of the sort key): 4096 x 2048."""
    from pldepth_b200 import ops
    H, W, K, n = 4096, 2048, 5, 2000
    g = torch.rand((1, H, W), device=cuda_device)
    g[0, -1, -1] = 2.0                                   # the last pixel is the deepest
    mask = torch.ones((1, H, W), device=cuda_device)
    pred = torch.randn((1, H, W, 1), device=cuda_device)
    loss, _, grad, rank, _, nv = ops.fused_step(mask, g, pred, K, n, seed=5)
    ops.check_status(cuda_device)
    assert int(nv[0]) == -(H * W)
    idx = rank[0, :, :, 0].long()
    assert int(idx.max()) < H * W and int(idx.max()) > (H * W) * 0.99        # draws reach the top of the range
    assert bool((torch.gather(g.reshape(-1), 0, idx.reshape(-1)).reshape(n, K) == rank[0, :, :, 1]).all())
    with pytest.raises(Exception):
        ops.fused_step(torch.ones((1, 4097, 2048), device=cuda_device), torch.rand((1, 4097, 2048), device=cuda_device),
                       torch.randn((1, 4097, 2048, 1), device=cuda_device), K, 10, seed=1)


@pytest.mark.parametrize("strategy", ["purely", "thresholded"])
def test_cuda_graph_replay_draws_fresh_lists(cuda_device, strategy):
    """FusedPLStep.capture: the Philox offset lives on the device, so each replay equals the eager step with
    the next offset (and differs from the previous replay)."""
    from pldepth_b200 import ops
    from pldepth_b200._lib import Context
    from pldepth_b200.step import FusedPLStep
    from tests.test_gpu_sampler import make_maps
    B, H, W, K, R = 2, 32, 32, 5, 300
    gt, mask = make_maps(H, W, H, W, 4, B)
    pred = np.random.RandomState(1).randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    st = FusedPLStep(K, R, seed=6, strategy=strategy)
    st.step_index = 10
    ctx = Context.current(cuda_device.index or 0)
    try:
        graph, buf = st.capture(gt_d, mask_d, pred_d)         # warm-up consumed offset 10
        seen = []
        for i in range(3):
            graph.replay()
            torch.cuda.synchronize()
            seen.append((buf["rankings"].clone(), buf["loss"].item()))
        ctx.device_offset(False)
        for i, (rank, loss) in enumerate(seen):
            if strategy == "purely":
                want = ops.fused_step(mask_d, gt_d, pred_d, K, R, seed=6, offset=11 + i)
                wrank, wloss = want[3], want[0].item()
            else:
                want = ops.fused_step_scored(mask_d, gt_d, pred_d, K, int(R * 1.5), R, strategy, seed=6, offset=11 + i)
                wrank, wloss = want["rankings"], want["loss"].item()
            assert torch.equal(rank, wrank) and loss == wloss
        assert not torch.equal(seen[0][0], seen[1][0])
    finally:
        ctx.device_offset(False)


@pytest.mark.parametrize("strategy,K,R", [("purely", 5, 400), ("information", 5, 3000), ("thresholded", 20, 7000)])
def test_cuda_graph_replay_without_rankings(cuda_device, strategy, K, R):
    """The training-step form (rankings not materialised) from a CUDA graph: holed masks run in bit-mask mode, the scored
    strategies through the sampled-window selection (more than 8192 candidates per image; over the stored keys for
    K > 16) -- every kernel of those chains is a programmatic dependent launch, which the capture must keep in order.
    Each replay equals the eager step with the next Philox offset."""
    from pldepth_b200._lib import Context
    from pldepth_b200.step import FusedPLStep
    from tests.test_gpu_sampler import make_maps
    B, H, W = 2, 48, 40
    gt, mask = make_maps(H, W, H, W, 9, B)
    mask[1] = 1.0                                   # one full mask, one holed
    pred = np.random.RandomState(2).randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    ctx = Context(cuda_device.index or 0)           # private context: the device-resident offset stays out of the way
    st = FusedPLStep(K, R, seed=8, strategy=strategy, emit_rankings=False, context=ctx)
    st.step_index = 20
    graph, buf = st.capture(gt_d, mask_d, pred_d)   # warm-up consumed offset 20
    seen = []
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        seen.append((buf["loss"].item(), buf["grad"].clone()))
    st.check(cuda_device)
    for i, (loss, grad) in enumerate(seen):
        ref = FusedPLStep(K, R, seed=8, strategy=strategy, emit_rankings=True)
        ref.step_index = 21 + i
        out = ref.run(gt_d, mask_d, pred_d)
        want_loss, want_grad, _ = lo.hourglass_nll(out["rankings"].cpu().numpy(), pred, B, K)
        assert_close(loss, want_loss, "replayed loss")
        assert_close(grad.cpu().numpy().reshape(B, -1), want_grad.reshape(B, -1), "replayed gradient")
    assert seen[0][0] != seen[1][0]


def test_pre_gathered_loss_and_standalone_gather(cuda_device):
    """NegativeLogLikelihoodLoss (nll_loss.py:10-29) and prepare_fully_fledged_loss_input (depth_utils.py:39-61)."""
    from pldepth_b200.depth_utils import get_depth_relation, prepare_fully_fledged_loss_input
    from pldepth_b200.losses import NegativeLogLikelihoodLoss
    B, H, W, K, R = 2, 16, 12, 6, 70
    y_true, pred = make_problem(B, H, W, K, R, 4, sorted_lists=False)
    yt, p = torch.from_numpy(y_true).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    sel, lab = prepare_fully_fledged_loss_input(yt, p, B, K)
    idx, want_lab = lo.split_rankings(y_true, B, K)
    want_sel = lo.gather_predictions(pred, idx, B, K)
    assert np.array_equal(sel.cpu().numpy(), want_sel) and np.array_equal(lab.cpu().numpy(), want_lab)
    logits = sel.clone().requires_grad_(True)
    val = NegativeLogLikelihoodLoss(K)(lab, logits)
    val.backward()
    nll, g = lo.listmle_per_list(want_lab, want_sel)
    assert_close(val.item(), nll.mean(), "pre-gathered loss")
    assert_close(logits.grad.cpu().numpy(), g / nll.shape[0], "pre-gathered gradient")
    assert get_depth_relation(1.0, 0.5) == 1 and get_depth_relation(0.5, 1.0) == -1 and get_depth_relation(2, 2) == 0
    assert get_depth_relation(1.02, 1.0, 0.03) == 0 and get_depth_relation(1.04, 1.0, 0.03) == 1
    assert get_depth_relation(1.0, 1.04, 0.03) == -1


@pytest.mark.parametrize("name,B,H,W,K,R", [("C3", 16, 448, 448, 50, 50000), ("C5-slice", 4, 1024, 768, 10, 1000000)])
def test_full_size_config3_and_config5_properties(cuda_device, name, B, H, W, K, R):
    """BASELINE config 3 at full size and 4 of the 32 images of one GPU's config-5 share: depth-descending
    lists that reproduce gt at their indices, loss == mean per-list NLL, gradient sums to ~0, and a random
    sample of lists reproduces the oracle's per-list NLL (long lists: reverse log-cumsum-exp of length 50)."""
    from pldepth_b200 import ops, synth
    rs = np.random.RandomState(1)
    base = synth.depth_map(H, W, 3000 + K)
    gt = np.stack([np.roll(base, 53 * b, axis=0) for b in range(B)])
    pred = rs.standard_normal((B, H, W, 1)).astype(np.float32)
    gt_d, pred_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(pred).to(cuda_device)
    mask_d = torch.ones((B, H, W), device=cuda_device)
    loss, loss_sum, grad, rank, per_list, nv = ops.fused_step(mask_d, gt_d, pred_d, K, R, seed=3, want_per_list=True)
    ops.check_status(cuda_device)
    d = rank[..., 1]
    assert bool((d[:, :, :-1] >= d[:, :, 1:]).all())
    idx = rank[..., 0].long()
    assert int(idx.min()) >= 0 and int(idx.max()) < H * W
    assert bool((torch.gather(gt_d.reshape(B, -1), 1, idx.reshape(B, -1)).reshape(B, R, K) == d).all())
    assert abs(loss.item() - per_list.double().mean().item()) <= 2e-6 * abs(loss.item())
    assert abs(loss_sum.item() - per_list.double().sum().item()) <= 2e-6 * abs(loss_sum.item())
    gsum, gabs = grad.double().sum().item(), grad.double().abs().sum().item()
    assert abs(gsum) <= 1e-5 * gabs
    pick = rs.randint(0, R, size=1500)
    b = B - 1
    _, _, want_pl = lo.hourglass_nll(rank[b, pick].cpu().numpy()[None], pred[b:b + 1], 1, K)
    assert_close(per_list.reshape(B, R)[b, pick].cpu().numpy(), want_pl, "%s per-list NLL at full size" % name)
    # uniform draws: every image's index histogram over 64 equal bands is flat within 6 sigma
    hist = torch.bincount((idx[b].reshape(-1) * 64) // (H * W), minlength=64).double()
    exp = R * K / 64.0
    assert float(((hist - exp).abs() / exp ** 0.5).max()) < 6.0


import glob as _glob
import os as _os

TFR_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "listmle_*.npz")))


@pytest.mark.skipif(not TFR_GOLDEN, reason="no tests/golden/listmle_*.npz yet (run tools/pin_tfranking.py where TF exists)")
@pytest.mark.parametrize("path", TFR_GOLDEN, ids=[_os.path.basename(p)[8:-4] for p in TFR_GOLDEN])
def test_cuda_loss_matches_tfranking_goldens(cuda_device, path):
    """The CUDA loss against outputs of the reference's own HourglassNegativeLogLikelihood (tools/pin_tfranking.py)."""
    from pldepth_b200.losses import HourglassNegativeLogLikelihood
    g = np.load(path)
    B, K = int(g["batch_size"]), int(g["ranking_size"])
    yt = torch.from_numpy(g["y_true"]).to(cuda_device)
    yp = torch.from_numpy(g["y_pred"]).to(cuda_device).requires_grad_(True)
    value = HourglassNegativeLogLikelihood(K, B)(yt, yp)
    value.backward()
    assert_close(value.item(), float(g["loss"]), "loss vs TF-Ranking")
    assert_close(yp.grad.cpu().numpy(), g["grad"], "gradient vs TF-Ranking")
