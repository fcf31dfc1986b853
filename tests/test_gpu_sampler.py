"""GPU parity tests, stage 1 (sampler), through the C ABI.  Bit-exact against the oracle and
against the reference's golden outputs (tests/golden)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mt19937_oracle as mto
from oracle import sampler_oracle as so
from tests import philox_model as pm
from tests.test_oracle_sampler import CLS2STRAT, load_case

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sampler_*.npz")))


def make_maps(H, W, Hm, Wm, seed, B=2, hole=True):
    rs = np.random.RandomState(seed)
    gt = np.stack([((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W) for _ in range(B)])
    mask = (rs.rand(B, Hm, Wm) > (0.25 if hole else -1)).astype(np.float32)
    return gt, mask


@pytest.mark.parametrize("H,W,Hm,Wm", [(24, 32, 24, 32), (24, 32, 12, 8), (30, 20, 7, 9), (448, 448, 448, 448),
                                       (65, 67, 65, 67)])
def test_mask_compact_matches_np_where(cuda_device, H, W, Hm, Wm):
    from pldepth_b200 import ops
    gt, mask = make_maps(H, W, Hm, Wm, 1, B=3)
    mask[1] = 1.0                       # a full mask
    mask[2, :, :] = 0
    mask[2, Hm - 1, Wm - 1] = 0.5       # a single valid pixel
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    vf, nv = vf.cpu().numpy(), nv.cpu().numpy()
    for b in range(3):
        want = so.valid_flat_indices(mask[b], (H, W))
        assert abs(nv[b]) == want.shape[0]
        if nv[b] < 0:       # identity shortcut: full mask at image resolution, row not materialised
            assert (H, W) == (Hm, Wm) and want.shape[0] == H * W
        else:
            assert np.array_equal(vf[b, :nv[b]], want)
    assert (nv[1] < 0) == ((H, W) == (Hm, Wm))


@pytest.mark.parametrize("K", [1, 2, 3, 5, 8, 10, 16, 17, 31, 32, 50, 64, 100, 128, 200, 256, 500, 512])
def test_philox_draws_and_rankings(cuda_device, K):
    """Philox mode: the draws equal the independent NumPy Philox model, and the rankings equal
    the oracle fed with those draws ("same fed indices" parity)."""
    from pldepth_b200 import ops
    H, W = 40, 50
    B = 2
    n = 257 if K <= 64 else 41
    gt, mask = make_maps(H, W, H, W, 100 + K, B)
    gt_d = torch.from_numpy(gt).to(cuda_device)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    seed, offset, base = 0x1234_5678_9ABC, 7 + (3 << 32), 5
    rank, sel = ops.sample_lists_philox(gt_d, vf, nv, K, n, seed, offset, base, want_sel=True)
    ops.check_status(cuda_device)
    rank, sel, nvh = rank.cpu().numpy(), sel.cpu().numpy(), nv.cpu().numpy()
    for b in range(B):
        want_sel = pm.draw_selection(seed, offset, base + b, n, K, abs(int(nvh[b])))
        assert np.array_equal(sel[b], want_sel)
        want = so.rankings_from_selection(sel[b].reshape(-1), so.valid_flat_indices(mask[b], (H, W)), gt[b], K)
        assert np.array_equal(rank[b], want)


@pytest.mark.parametrize("K", [5, 20])
def test_philox_redraw_path_on_large_maps(cuda_device, K):
    """M = 3 M valid pixels (2^32 mod M = 1 967 296): Lemire's rejection fires about once per 2200
    draws, so the redraw stream (word block 0x8000 | ...) is exercised and must match the model."""
    from pldepth_b200 import ops
    H, W = 1500, 2000
    n = 12000 if K == 5 else 3000
    rs = np.random.RandomState(1)
    gt = rs.rand(1, H, W).astype(np.float32)
    mask = np.ones((1, H, W), np.float32)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    M = abs(int(nv[0].item()))
    assert M == H * W
    _, sel = ops.sample_lists_philox(torch.from_numpy(gt).to(cuda_device), vf, nv, K, n, 99, 0, 0, want_sel=True,
                                     want_rankings=False)
    want = pm.draw_selection(99, 0, 0, n, K, M)
    thresh = ((1 << 32) - M) % M
    low = (pm.philox4x32_10(np.arange(n, dtype=np.uint32), 0, 0, 0, 99, 0)[0].astype(np.uint64) * np.uint64(M)) & np.uint64(0xFFFFFFFF)
    assert thresh == 1967296 and (low < thresh).sum() > 0        # the redraw path really runs
    assert np.array_equal(sel[0].cpu().numpy(), want)


def test_philox_is_independent_of_batch_split(cuda_device):
    from pldepth_b200 import ops
    H, W, K, n = 32, 32, 5, 300
    gt, mask = make_maps(H, W, H, W, 3, B=4)
    gt_d, mask_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    vf, nv = ops.mask_compact(mask_d, H, W)
    full, _ = ops.sample_lists_philox(gt_d, vf, nv, K, n, 42, 1, 0)
    vf2, nv2 = ops.mask_compact(mask_d[2:], H, W)
    part, _ = ops.sample_lists_philox(gt_d[2:], vf2, nv2, K, n, 42, 1, 2)
    assert torch.equal(full[2:], part)


@pytest.mark.parametrize("K", [4, 5, 16, 20, 64])
def test_ties_later_draw_first(cuda_device, K):
    """Heavily tied depths: same rule as the oracle (reversed stable argsort)."""
    from pldepth_b200 import ops
    H, W, n = 16, 16, 200
    rs = np.random.RandomState(K)
    gt = (rs.randint(0, 4, size=(1, H, W)) / 4).astype(np.float32)
    mask = np.ones((1, H, W), np.float32)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    sel = rs.randint(0, H * W, size=(1, n, K)).astype(np.int32)
    rank = ops.sample_lists_fed(torch.from_numpy(gt).to(cuda_device), vf, nv, K, torch.from_numpy(sel).to(cuda_device))
    want = so.rankings_from_selection(sel.reshape(-1), np.arange(H * W), gt[0], K)
    assert np.array_equal(rank[0].cpu().numpy(), want)


def test_device_mt19937_equals_numpy(cuda_device):
    from pldepth_b200 import ops
    for seed in (0, 3, 2 ** 32 - 1):
        state, pos = ops.mt19937_init(seed, cuda_device)
        assert np.array_equal(state.cpu().numpy().view(np.uint32), mto.init_genrand(seed))
        rs = np.random.RandomState(seed)
        for n in (1, 623, 624, 625, 5000):
            got = ops.mt19937_generate(state, pos, n).cpu().numpy().view(np.uint32)
            want = rs.randint(0, 2 ** 32, size=n, dtype=np.uint32)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
@pytest.mark.parametrize("rng", ["numpy", "mt19937"])
def test_sampler_classes_reproduce_reference_goldens(cuda_device, path, rng):
    """The drop-in classes, fed like the reference (NumPy image/mask/gt), reproduce the
    reference's own outputs bit for bit, and leave np.random in the same state."""
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    c = load_case(path)
    cls = {v: k for k, v in CLS2STRAT.items()}[c["strategy"]]
    strat = getattr(sampling, cls)(ModelParameters(ranking_size=c["K"]), rng=rng, seed=c["seed"])
    H, W = c["gt"].shape
    image = np.zeros((H, W, 3), np.float32)
    np.random.seed(c["seed"])
    st0 = np.random.get_state()
    if c["factor"] < 0:
        out = strat.sample_masked_point_batch(image, c["mask"], c["gt"], c["R"])
    elif c["strategy"] == "purely":
        out, dists = strat.sample_masked_rankings(image, c["mask"], c["gt"], c["R"], c["factor"])
        assert dists.shape == (out.shape[0],)
    else:
        out = strat.sample_masked_point_batch(image, c["mask"], c["gt"], c["R"], c["factor"])
    assert out.dtype == np.float32
    assert np.array_equal(out, c["rankings"])
    if rng == "numpy":
        st1 = np.random.get_state()
        rs = np.random.RandomState()
        rs.set_state(st0)
        rs.randint(0, 2 ** 32, size=c["consumed"], dtype=np.uint32)
        st2 = rs.get_state()
        assert np.array_equal(st1[1], st2[1]) and st1[2] == st2[2]


@pytest.mark.parametrize("strategy", ["masked", "thresholded", "information"])
@pytest.mark.parametrize("promotion", ["nep50", "legacy"])
@pytest.mark.parametrize("K,n,R", [(2, 500, 333), (5, 500, 333), (7, 500, 333), (8, 500, 333), (9, 500, 333),
                                   (20, 500, 333), (130, 500, 333), (300, 500, 333),
                                   # more than 8192 candidates per image: segmented radix sort instead of the
                                   # shared-memory sort
                                   (5, 9000, 4100), (9, 8193, 8193), (20, 12000, 1)])
def test_scores_and_selection_match_oracle(cuda_device, strategy, promotion, K, n, R):
    from pldepth_b200 import ops
    from tests.golden.make_golden import near_threshold_gt
    H, W, B = 48, 40, 3
    rs = np.random.RandomState(K)
    gt = np.stack([near_threshold_gt(H, W, 7 * K + b) for b in range(B)])
    mask = np.ones((B, H, W), np.float32)
    sel = rs.randint(0, H * W, size=(B, n, K)).astype(np.int32)
    gt_d = torch.from_numpy(gt).to(cuda_device)
    vf, nv = ops.mask_compact(torch.from_numpy(mask).to(cuda_device), H, W)
    cand = ops.sample_lists_fed(gt_d, vf, nv, K, torch.from_numpy(sel).to(cuda_device))
    mm = ops.gt_minmax(gt_d) if strategy == "information" else None
    scores = ops.score_lists(cand, strategy, 0.03, -1000, promotion, mm)
    top, order = ops.select_top(scores, cand, R, want_order=True)
    cand_h, scores_h, top_h, order_h = cand.cpu().numpy(), scores.cpu().numpy(), top.cpu().numpy(), order.cpu().numpy()
    for b in range(B):
        if strategy == "information":
            want = so.score_information(cand_h[b], gt[b], 0.03, -1000, promotion)
        else:
            want = so.score_adjacent_differences(cand_h[b], 0.03 if strategy == "thresholded" else None, -1000,
                                                 promotion)
        assert np.array_equal(scores_h[b], want), (np.abs(scores_h[b] - want).max())
        want_top, want_order = so.select_top(cand_h[b], want, R)
        assert np.array_equal(order_h[b], want_order)
        assert np.array_equal(top_h[b], want_top)


def test_empty_mask_raises_like_randint0(cuda_device):
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    strat = sampling.PurelyMaskedRandomSamplingStrategy(ModelParameters(ranking_size=3), rng="philox")
    image = np.zeros((8, 8, 3), np.float32)
    with pytest.raises(ValueError):
        strat.sample_masked_point_batch(image, np.zeros((8, 8), np.float32), np.ones((8, 8), np.float32), 10)


def test_single_valid_pixel_consumes_no_words(cuda_device):
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    strat = sampling.PurelyMaskedRandomSamplingStrategy(ModelParameters(ranking_size=3), rng="numpy")
    mask = np.zeros((8, 8), np.float32)
    mask[2, 5] = 1
    gt = np.random.RandomState(0).rand(8, 8).astype(np.float32)
    np.random.seed(4)
    st0 = np.random.get_state()
    out = strat.sample_masked_point_batch(np.zeros((8, 8, 3)), mask, gt, 10)
    st1 = np.random.get_state()
    assert out.shape == (8, 3, 2) and (out[:, :, 0] == 21).all() and (out[:, :, 1] == gt[2, 5]).all()
    assert np.array_equal(st0[1], st1[1]) and st0[2] == st1[2]


def test_provider_mirror_per_image_and_batched(cuda_device):
    """HourglassLargeScaleDataProvider.sample_rankings keeps the reference's per-image contract and
    reproduces a golden case; the batched device path returns the (B, R, K, 2) tensor the reference's
    map + batch would."""
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    from pldepth_b200.provider import HourglassLargeScaleDataProvider
    c = load_case([p for p in GOLDEN if "thresholded_k5" in p][0])
    mp = ModelParameters(ranking_size=c["K"], rankings_per_image=c["R"], batch_size=2, val_rankings_per_img=7)
    mp.set_parameter("sampling_strategy", sampling.ThresholdedMaskedRandomSamplingStrategy(mp))
    prov = HourglassLargeScaleDataProvider(mp, None, None)
    H, W = c["gt"].shape
    image = np.zeros((H, W, 3), np.float64)
    np.random.seed(c["seed"])
    img_out, rank = prov.sample_rankings(image, c["mask"], c["gt"])
    assert img_out.dtype == np.float32 and rank.dtype == np.float32
    assert np.array_equal(rank, c["rankings"])
    val = prov.generate_validation_rankings([(image, c["mask"], c["gt"])] * 3)
    assert val.shape == (3, 7, c["K"], 2) and (np.diff(val[..., 1], axis=-1) <= 0).all()
    # batched, philox
    mp2 = ModelParameters(ranking_size=5, rankings_per_image=64, batch_size=3)
    mp2.set_parameter("sampling_strategy", sampling.InformationScoreBasedSampling(mp2, rng="philox", seed=3))
    prov2 = HourglassLargeScaleDataProvider(mp2, augmentation=True)
    gts = [c["gt"]] * 7
    masks = [c["mask"]] * 7
    images = [image] * 7
    it = prov2.iterate_train_batches(images, masks, gts, cuda_device, repeat=False)
    batches = list(it)
    assert len(batches) == 2                                # drop_remainder: 7 // 3
    img_b, y_true = batches[0]
    assert tuple(img_b.shape) == (3, H, W, 3) and tuple(y_true.shape) == (3, 64, 5, 2)
    d = y_true[..., 1]
    assert bool((d[:, :, :-1] >= d[:, :, 1:]).all())


@pytest.mark.parametrize("strategy", ["masked", "thresholded", "information"])
@pytest.mark.parametrize("promotion", ["nep50", "legacy"])
@pytest.mark.parametrize("K,geometry", [(2, "full"), (5, "holes"), (5, "ties"), (8, "scaled"), (9, "full"), (16, "holes"),
                                        (17, "full"), (20, "holes"), (33, "ties"), (50, "scaled"), (64, "full"),
                                        (100, "ties"), (128, "holes"), (130, "holes"), (300, "full")])
@pytest.mark.parametrize("n,R", [(700, 333), (8192, 8192), (9000, 4100)], ids=["smem", "smem-max", "radix"])
def test_scored_step_equals_staged_pipeline(cuda_device, strategy, promotion, K, geometry, n, R):
    """pld_fused_step_scored (score-only pass + radix top-R + redraw) returns the same kept candidates, in
    the same order, with the same rankings as the staged sample -> score -> select_top pipeline (which the
    tests above pin to the oracle), and its loss / gradient match the oracle."""
    from oracle import listmle_oracle as lo
    from pldepth_b200 import ops
    from tests.golden.make_golden import near_threshold_gt
    B, H, W = 3, 40, 36
    Hm, Wm = (20, 12) if geometry == "scaled" else (H, W)
    rs = np.random.RandomState(K)
    if geometry == "ties":       # few distinct depths: massive score ties -> tie rule and a big boundary bucket
        gt = (rs.randint(0, 6, size=(B, H, W)) / 8 + 0.1).astype(np.float32)
    else:
        gt = np.stack([near_threshold_gt(H, W, 11 * K + b) for b in range(B)])
    mask = np.ones((B, Hm, Wm), np.float32)
    if geometry in ("holes", "scaled"):
        mask = (rs.rand(B, Hm, Wm) > 0.3).astype(np.float32)
        if geometry == "holes":
            mask[1] = 1.0
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    out = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, promotion, seed=4, offset=2,
                                image_base=1, want_order=True, want_per_list=True)
    vf, nv = ops.mask_compact(mask_d, H, W)
    cand, _ = ops.sample_lists_philox(gt_d, vf, nv, K, n, 4, 2, 1)
    mm = ops.gt_minmax(gt_d) if strategy == "information" else None
    scores = ops.score_lists(cand, strategy, 0.03, -1000, promotion, mm)
    top, order = ops.select_top(scores, cand, R, want_order=True)
    ops.check_status(cuda_device)
    assert torch.equal(out["order"], order)
    assert torch.equal(out["rankings"], top)
    want_loss, want_grad, want_pl = lo.hourglass_nll(top.cpu().numpy(), pred, B, K)
    assert abs(out["loss"].item() - want_loss) <= 1e-5 * abs(want_loss)
    err = np.abs(out["grad"].cpu().numpy() - want_grad).max() / np.abs(want_grad).max()
    assert err <= 1e-5
    # sampler-only variant (no pred): same rankings
    out2 = ops.fused_step_scored(mask_d, gt_d, None, K, n, R, strategy, 0.03, -1000, promotion, seed=4, offset=2,
                                 image_base=1)
    assert torch.equal(out2["rankings"], top) and out2["loss"] is None
    # rankings not materialised: exact radix selection instead of the sort -- same kept SET, same loss / gradient
    out3 = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, promotion, seed=4, offset=2,
                                 image_base=1, want_rankings=False, want_order=True)
    assert out3["rankings"] is None
    assert torch.equal(torch.sort(out3["order"], dim=1).values, torch.sort(order, dim=1).values)
    assert abs(out3["loss"].item() - want_loss) <= 1e-5 * abs(want_loss)
    assert np.abs(out3["grad"].cpu().numpy() - want_grad).max() / np.abs(want_grad).max() <= 1e-5


def test_strategy_classes_use_fast_path_consistently(cuda_device):
    """sample_batch in Philox mode (fast scored path) == draw_candidates + score + select (staged)."""
    from pldepth_b200 import ops, sampling
    from pldepth_b200.models_meta import ModelParameters
    rs = np.random.RandomState(0)
    B, H, W, K, R = 2, 32, 32, 5, 100
    gt = rs.rand(B, H, W).astype(np.float32)
    mask = (rs.rand(B, H, W) > 0.2).astype(np.float32)
    gt_d, mask_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    for cls in (sampling.MaskedRandomSamplingStrategy, sampling.ThresholdedMaskedRandomSamplingStrategy,
                sampling.InformationScoreBasedSampling):
        a = cls(ModelParameters(ranking_size=K), rng="philox", seed=9)
        b = cls(ModelParameters(ranking_size=K), rng="philox", seed=9)
        fast = a.sample_batch(gt_d, mask_d, R)
        n = int(R * b._default_factor)
        cand = b.draw_candidates(gt_d, mask_d, n)
        staged, _ = ops.select_top(b.score_candidates(cand, gt_d), cand, R)
        assert torch.equal(fast, staged)


def test_samplers_are_thread_safe(cuda_device):
    """The reference's sampler is called concurrently from tf.data worker threads
    (hourglass_provider.py:55-58, num_parallel_calls=AUTOTUNE).  Each thread gets its own pld_ctx; results
    must be valid and, in Philox mode, independent of the interleaving."""
    from concurrent.futures import ThreadPoolExecutor
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    rs = np.random.RandomState(0)
    H, W, K, R = 48, 48, 5, 200
    gts = [((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W) for _ in range(8)]
    masks = [(rs.rand(H, W) > 0.2).astype(np.float32) for _ in range(8)]
    image = np.zeros((H, W, 3), np.float32)

    def work(i):
        s = sampling.ThresholdedMaskedRandomSamplingStrategy(ModelParameters(ranking_size=K), rng="philox", seed=100 + i)
        outs = [s.sample_masked_point_batch(image, masks[i], gts[i], R) for _ in range(3)]
        return outs

    with ThreadPoolExecutor(max_workers=8) as ex:
        par = list(ex.map(work, range(8)))
    seq = [work(i) for i in range(8)]
    for i in range(8):
        for a, b in zip(par[i], seq[i]):
            assert a.shape == (R, K, 2) and np.array_equal(a, b)
            assert (np.diff(a[:, :, 1], axis=1) <= 0).all()
            assert (masks[i].reshape(-1)[a[:, :, 0].astype(np.int64)] > 0).all()
    # numpy-compat mode from several threads: serialised on the global RNG, every result well formed
    def work_np(i):
        s = sampling.PurelyMaskedRandomSamplingStrategy(ModelParameters(ranking_size=K), rng="numpy")
        return s.sample_masked_point_batch(image, masks[i], gts[i], R)
    with ThreadPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(work_np, range(8)))
    for i, a in enumerate(res):
        assert a.shape == (int(R * 0.8), K, 2)
        assert (masks[i].reshape(-1)[a[:, :, 0].astype(np.int64)] > 0).all()


@pytest.mark.parametrize("strategy,cls", [("thresholded", "ThresholdedMaskedRandomSamplingStrategy"),
                                          ("information", "InformationScoreBasedSampling")])
def test_numpy_compat_mode_at_scale(cuda_device, strategy, cls):
    """np.random-compatible mode with hundreds of thousands of draws per image (many compaction tiles,
    multi-block scans, long raw windows): still bit-exact against the oracle run on the same NumPy stream."""
    from pldepth_b200 import sampling
    from pldepth_b200.models_meta import ModelParameters
    H = W = 160
    K, R = 5, 20000
    rs = np.random.RandomState(8)
    gt = ((rs.permutation(H * W) + 0.5) / (H * W)).astype(np.float32).reshape(H, W)
    mask = (rs.rand(H, W) > 0.35).astype(np.float32)
    image = np.zeros((H, W, 3), np.float32)
    strat = getattr(sampling, cls)(ModelParameters(ranking_size=K), rng="numpy")
    np.random.seed(77)
    got = strat.sample_masked_point_batch(image, mask, gt, R)
    st_after = np.random.get_state()
    np.random.seed(77)
    want, _, scores = so.sample_masked_point_batch(strategy, (H, W), mask, gt, R, K)
    st_want = np.random.get_state()
    assert got.shape == want.shape == (R, K, 2)
    if np.unique(scores).size == scores.size:
        assert np.array_equal(got, want)
    else:   # tied scores: same rule on both sides, still identical
        assert np.array_equal(got, want)
    assert np.array_equal(st_after[1], st_want[1]) and st_after[2] == st_want[2]


def _want_order(scores, R):
    return np.argsort(scores, kind="stable")[::-1][:R].astype(np.int32)


@pytest.mark.parametrize("kind", ["all-equal", "two-values", "dense-f32", "wide-f64", "clustered"])
@pytest.mark.parametrize("n,R", [(20000, 7000), (300000, 299999)])
def test_top_selection_on_adversarial_scores(cuda_device, kind, n, R):
    """The segmented radix sort behind pld_select_top (more than 8192 candidates per image) on awkward score
    distributions: all equal, two values (two-valued digit bytes), one binade, 600 decades, two far-apart clusters.
    Result = reversed stable argsort (sampling.py:169, 208, 239)."""
    from pldepth_b200 import ops
    rs = np.random.RandomState(n % 1000 + len(kind))
    B = 2
    if kind == "all-equal":
        sc = np.full((B, n), -3.25)
    elif kind == "two-values":
        sc = rs.choice([-1000.5, 0.125], size=(B, n))
    elif kind == "dense-f32":
        sc = rs.rand(B, n).astype(np.float32).astype(np.float64) * 0.5 + 0.5
    elif kind == "wide-f64":
        sc = rs.standard_normal((B, n)) * 10.0 ** rs.randint(-300, 300, size=(B, n))
    else:
        sc = np.where(rs.rand(B, n) < 0.5, -1000.0, 0.0) - rs.rand(B, n) * 1e-3
    cand = torch.arange(B * n, dtype=torch.float32, device=cuda_device).reshape(B, n, 1, 1).repeat(1, 1, 1, 2)
    top, order = ops.select_top(torch.from_numpy(sc).to(cuda_device), cand, R, want_order=True)
    ops.check_status(cuda_device)
    order_h = order.cpu().numpy()
    for b in range(B):
        assert np.array_equal(order_h[b], _want_order(sc[b], R))
    assert torch.equal(top[..., 0, 0], (order.long() + torch.arange(B, device=cuda_device)[:, None] * n).float())


def test_top_selection_very_long_segment(cuda_device):
    """1.6 M candidates in one image (hundreds of tiles per image in the segmented sort)."""
    from pldepth_b200 import ops
    n, R = 1_600_000, 1_000_000
    rs = np.random.RandomState(3)
    sc = np.round(rs.standard_normal((1, n)), 3)          # plenty of ties
    cand = torch.zeros((1, n, 1, 2), dtype=torch.float32, device=cuda_device)
    _, order = ops.select_top(torch.from_numpy(sc).to(cuda_device), cand, R, want_order=True)
    ops.check_status(cuda_device)
    assert np.array_equal(order.cpu().numpy()[0], _want_order(sc[0], R))


@pytest.mark.parametrize("strategy", ["thresholded", "information"])
def test_scored_step_at_config2_scale(cuda_device, strategy):
    """BASELINE config-2 shapes per image (448 x 448, K = 5, R = 100 000; 150 000 / 500 000 candidates): the one-call
    scored step still equals the staged sample -> score -> select pipeline bit for bit, and the unordered selection
    keeps the same set.  Exercises many selection tiles, multi-tile radix passes and the tie-free shortcut."""
    from pldepth_b200 import ops, synth
    B, H, W, K, R = 2, 448, 448, 5, 100000
    n = int(R * (1.5 if strategy == "thresholded" else 5))
    gt = np.stack([synth.depth_map(H, W, 40 + b) for b in range(B)])
    if strategy == "thresholded":
        gt[1] = np.round(gt[1] * 64) / 64 + 0.01          # an image whose scores tie massively
    mask = np.stack([synth.valid_mask(H, W, 50 + b, 0.1 * b) for b in range(B)])
    gt_d, mask_d = torch.from_numpy(gt).to(cuda_device), torch.from_numpy(mask).to(cuda_device)
    pred_d = torch.randn((B, H, W, 1), device=cuda_device)
    out = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, "nep50", seed=9, offset=1,
                                want_order=True)
    vf, nv = ops.mask_compact(mask_d, H, W)
    cand, _ = ops.sample_lists_philox(gt_d, vf, nv, K, n, 9, 1, 0)
    mm = ops.gt_minmax(gt_d) if strategy == "information" else None
    scores = ops.score_lists(cand, strategy, 0.03, -1000, "nep50", mm)
    top, order = ops.select_top(scores, cand, R, want_order=True)
    ops.check_status(cuda_device)
    assert torch.equal(out["order"], order) and torch.equal(out["rankings"], top)
    # the staged order itself against NumPy's reversed stable argsort
    sc_h = scores.cpu().numpy()
    for b in range(B):
        assert np.array_equal(order[b].cpu().numpy(), _want_order(sc_h[b], R))
    out3 = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, 0.03, -1000, "nep50", seed=9, offset=1,
                                 want_rankings=False, want_order=True)
    assert torch.equal(torch.sort(out3["order"], dim=1).values, torch.sort(order, dim=1).values)
    assert abs(out3["loss"].item() - out["loss"].item()) <= 1e-6 * abs(out["loss"].item())


@pytest.mark.parametrize("strategy,K,n,R,kind", [
    ("thresholded", 5, 9000, 6000, "smooth"),
    ("information", 5, 40000, 8000, "smooth"),
    ("masked", 3, 20000, 19999, "ties"),        # a handful of distinct scores: the boundary bucket is huge
    ("information", 8, 12000, 1, "smooth"),
    ("thresholded", 2, 30000, 30000, "ties"),   # R == n: everything is kept
    ("masked", 16, 10000, 2500, "const"),       # one distinct score: the cut is decided by candidate ids alone
    # ranking_size > 16: the window selection runs over the key array the long-list scoring passes store
    ("thresholded", 20, 9000, 6000, "smooth"),
    ("information", 50, 12000, 3000, "smooth"),
    ("masked", 33, 10000, 2500, "const"),
    ("thresholded", 64, 8200, 8199, "ties"),
    ("information", 130, 9000, 100, "ties"),
])
@pytest.mark.parametrize("z", [None, "0", "0.5"])
def test_sampled_window_selection_keeps_the_exact_set(cuda_device, strategy, K, n, R, kind, z):
    """Rankings not materialised and more than 8192 candidates per image: the top-R set comes from the sampled-window
    selection (pld_pilot.cu).  It must be exactly the set the ordered pipeline keeps -- also when the window misses
    (PLD_PILOT_Z = 0 / 0.5 shrink it until the exact fallback runs for most images) and for holed masks."""
    import os
    from pldepth_b200 import ops, synth
    B, H, W = 3, 96, 80
    rs = np.random.RandomState(n + R)
    if kind == "smooth":
        gt = np.stack([synth.depth_map(H, W, 70 + b) for b in range(B)])
    elif kind == "ties":
        gt = (rs.randint(0, 5, size=(B, H, W)) / 8 + 0.1).astype(np.float32)
    else:
        gt = np.full((B, H, W), 0.5, np.float32)
    mask = np.stack([synth.valid_mask(H, W, 80 + b, 0.15 * b) for b in range(B)])
    pred = rs.randn(B, H, W, 1).astype(np.float32)
    gt_d, mask_d, pred_d = (torch.from_numpy(x).to(cuda_device) for x in (gt, mask, pred))
    ref = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, seed=21, offset=5, want_order=True)
    old = os.environ.get("PLD_PILOT_Z")
    try:
        if z is not None:
            os.environ["PLD_PILOT_Z"] = z
        out = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, seed=21, offset=5, want_rankings=False,
                                    want_order=True)
        ops.check_status(cuda_device)
    finally:
        if old is None:
            os.environ.pop("PLD_PILOT_Z", None)
        else:
            os.environ["PLD_PILOT_Z"] = old
    assert torch.equal(torch.sort(out["order"], dim=1).values, torch.sort(ref["order"], dim=1).values)
    assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-5 * abs(ref["loss"].item())
    g, g_ref = out["grad"].cpu().numpy(), ref["grad"].cpu().numpy()
    assert np.abs(g - g_ref).max() <= 1e-5 * np.abs(g_ref).max()
    # the order-preserving selection (PLD_NO_PILOT) keeps the same set as well
    os.environ["PLD_NO_PILOT"] = "1"
    try:
        out_old = ops.fused_step_scored(mask_d, gt_d, pred_d, K, n, R, strategy, seed=21, offset=5, want_rankings=False,
                                        want_order=True)
    finally:
        os.environ.pop("PLD_NO_PILOT", None)
    assert torch.equal(out_old["order"], torch.sort(ref["order"], dim=1).values)
