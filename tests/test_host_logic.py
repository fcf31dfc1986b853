"""CPU: host-side mirrors of the reference interface (no GPU needed)."""
import numpy as np
import pytest

from pldepth_b200 import losses, sampling
from pldepth_b200.models_meta import ModelParameters


def test_model_parameters_bag():
    mp = ModelParameters(ranking_size=5)
    assert mp.get_parameter("ranking_size") == 5
    assert mp.get_parameter("missing") is None and mp.get_parameter("missing", 3) == 3
    mp.set_parameter("batch_size", 4)
    dup = mp.duplicate()
    dup.set_parameter("batch_size", 8)
    assert mp.get_parameter("batch_size") == 4
    assert "ranking_size_5" in mp.get_parameter_string()


@pytest.mark.parametrize("cls,factor", [("PurelyMaskedRandomSamplingStrategy", 0.8),
                                        ("MaskedRandomSamplingStrategy", 1.5),
                                        ("ThresholdedMaskedRandomSamplingStrategy", 1.5),
                                        ("InformationScoreBasedSampling", 5)])
def test_sampler_construction_mirrors_reference(cls, factor):
    mp = ModelParameters(ranking_size=7, downscaling_factor=2)
    s = getattr(sampling, cls)(mp)
    assert s.num_points_per_sample == 7
    assert s.downscaling_factor == 2
    assert s.threshold == 0.03
    assert s._default_factor == factor
    assert cls in str(s) and "num_points_per_sample=7" in str(s)
    s.num_points_per_sample = 9
    assert s.num_points_per_sample == 9
    assert s.determine_x_y_scales(np.zeros((8, 6, 3)), np.zeros((4, 3))) == (2.0, 2.0)


def test_thresholded_positional_threshold_like_the_provider():
    # hourglass_provider.py:21 passes the threshold positionally
    s = sampling.ThresholdedMaskedRandomSamplingStrategy(ModelParameters(ranking_size=3), 0.05)
    assert s.threshold == 0.05 and s.equality_penalty == -1000
    with pytest.raises(ValueError):
        sampling.PurelyMaskedRandomSamplingStrategy(ModelParameters(ranking_size=3), rng="xorshift")


def test_unmasked_sampling_is_out_of_scope():
    s = sampling.RandomSamplingStrategy(ModelParameters(ranking_size=3))
    with pytest.raises(NotImplementedError):
        s.sample_points(None, None)


def test_loss_constructor_contract():
    fn = losses.HourglassNegativeLogLikelihood(ranking_size=5, batch_size=4, debug=False)
    assert fn.reduction == "auto" and fn.get_config()["ranking_size"] == 5
    with pytest.raises(NotImplementedError):
        losses.HourglassNegativeLogLikelihood(5, 4, lambda_weight=1)
    with pytest.raises(ValueError):
        losses.HourglassNegativeLogLikelihood(5, 4, reduction="median")
    with pytest.raises(ValueError):
        losses.HourglassNegativeLogLikelihood(0, 4)

    class FakeReduction(object):        # tf.losses.Reduction members carry a .name
        name = "SUM_OVER_BATCH_SIZE"
    assert losses.HourglassNegativeLogLikelihood(5, 4, reduction=FakeReduction()).reduction == "sum_over_batch_size"


def test_no_cpu_fallback():
    import torch
    from pldepth_b200 import ops
    from pldepth_b200._lib import PLDError
    with pytest.raises(PLDError):
        ops.as_cuda(torch.zeros(3), torch.float32, "x")


def test_algorithmic_bytes_match_the_survey_figures():
    """fused op = 16 B/pixel + 8 B/point (+ loss): 72.1 B/list at config 2, 464 at config 3,
    92.6 at config 5 -- the numerator of bench.py's roofline."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for (B, H, W, K, R), want in (((32, 448, 448, 5, 100000), 72.1), ((16, 448, 448, 50, 50000), 464.0),
                                  ((256, 1024, 768, 10, 1000000), 92.6)):
        L = B * R
        got = bench.algorithmic_bytes(B, H * W, L, K) / L
        assert abs(got - want) < 0.3, (got, want)     # the survey rounds to three digits


def test_comparator_networks_sort_and_match_the_committed_header():
    """pld_oem_networks.cuh (ordering networks of the thread-per-list scoring pass) is generated: the generator's
    networks sort (random and 0-1 inputs, descending) and the committed header is exactly what it writes."""
    import importlib.util
    import os
    import random
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_oem_network", os.path.join(root, "tools", "gen_oem_network.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rnd = random.Random(5)
    header = open(os.path.join(root, "pldepth_b200", "csrc", "pld_oem_networks.cuh")).read()
    for n in gen.SIZES:
        ces = gen.network(n)
        assert all(0 <= a < b < n for a, b in ces)
        for hi in (1, 7, 1 << 31):
            for _ in range(300):
                v = [rnd.randint(0, hi) for _ in range(n)]
                w = v[:]
                for a, b in ces:
                    if w[a] < w[b]:
                        w[a], w[b] = w[b], w[a]
                assert w == sorted(v, reverse=True)
        assert "// N = %d: %d comparators" % (n, len(ces)) in header
        block = header[header.index("struct OemNetwork<%d>" % n):]
        block = block[:block.index("#undef PLD_OEM_CE")]
        listed = [(int(a), int(b)) for a, b in re.findall(r"PLD_OEM_CE\((\d+), (\d+)\)", block)]
        assert listed == ces, "pld_oem_networks.cuh is stale: run tools/gen_oem_network.py"


def test_numpy_pairwise_sum_is_eight_strided_accumulators():
    """What the scoring kernels rely on (pld_score.cuh: np_pairwise_sum; pld_score_reg.cu, pld_lists_tab.cu): NumPy's
    float32 sum of 8 <= n <= 128 contiguous terms equals eight strided accumulators r[j] = a[j] + a[8 + j] + ...
    combined as ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), then the tail added in order; n < 8 is sequential."""
    import numpy as np
    rs = np.random.RandomState(11)
    f = np.float32
    for n in list(range(1, 40)) + [50, 56, 64, 100, 127, 128]:
        for _ in range(20):
            a = (rs.rand(n) * 10.0 ** rs.randint(-6, 3)).astype(np.float32)
            if n < 8:
                want = f(0)
                for x in a:
                    want = f(want + x)
            else:
                r = [f(a[j]) for j in range(8)]
                n8 = n - n % 8
                for i in range(8, n8, 8):
                    for j in range(8):
                        r[j] = f(r[j] + a[i + j])
                want = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
                for i in range(n8, n):
                    want = f(want + a[i])
            got = np.add.reduce(a)
            assert got.dtype == np.float32 and got.tobytes() == np.float32(want).tobytes(), (n, got, want)


def test_depth_relation_margin_test_never_disagrees_with_the_division():
    """ScoreCfg::c_in / c_out (pld_score.cuh: relation_equal): for a >= c > 0 the scoring kernels decide
    get_depth_relation (depth_utils.py:5-21) without dividing whenever a < c * thr_hi (1 - 2^-20) ("equal": the rounded
    ratio is below thr_hi, and >= 1 > thr_lo) or a > c * thr_hi (1 + 2^-20) ("not equal").  Emulated here in float32:
    wherever the margin test decides, the exact rounded division agrees -- including ratios within a few ulps of the
    margins and of the threshold itself."""
    import numpy as np
    f = np.float32
    rs = np.random.RandomState(3)
    for threshold in (0.03, 0.25, 1e-3, 2.0):
        thr_hi, thr_lo = 1.0 + threshold, 1.0 / (1.0 + threshold)
        thr_hi_f, thr_lo_f = f(thr_hi), f(thr_lo)
        c_in, c_out = f(thr_hi * (1.0 - 2.0 ** -20)), f(thr_hi * (1.0 + 2.0 ** -20))
        assert c_in < thr_hi_f < c_out and thr_lo_f < 1.0
        c = (rs.rand(400000) * 10.0 ** rs.randint(-6, 4, size=400000)).astype(np.float32) + f(1e-10)
        # ratios spread over [1, 2 thr_hi], plus clusters hugging the threshold and both margins
        ratio = np.concatenate([1.0 + rs.rand(100000) * (2 * thr_hi - 1.0),
                                thr_hi * (1.0 + (rs.rand(100000) - 0.5) * 2.0 ** -18),
                                thr_hi * (1.0 - 2.0 ** -20) * (1.0 + (rs.rand(100000) - 0.5) * 2.0 ** -21),
                                thr_hi * (1.0 + 2.0 ** -20) * (1.0 + (rs.rand(100000) - 0.5) * 2.0 ** -21)])
        a = (c.astype(np.float64) * ratio).astype(np.float32)
        keep = a >= c
        a, c = a[keep], c[keep]
        inside = a < (c * c_in)                       # float32 products, as __fmul_rn
        outside = a > (c * c_out)
        r = a / c                                     # float32 division, round to nearest
        equal = ~(r >= thr_hi_f) & ~(r <= thr_lo_f)
        assert not (inside & outside).any()
        assert equal[inside].all()
        assert not equal[outside].any()
        assert (inside | outside).mean() > 0.5        # the exact path is the exception
