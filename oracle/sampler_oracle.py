"""NumPy restatement of the reference ranking samplers -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pldepth/data/sampling.py (cited per function as ``sampling.py:L``)
and pldepth/data/depth_utils.py:5-21 (``get_depth_relation``).  Pinned bit-exactly against
the unmodified reference file run in the build container (tests/test_oracle_sampler.py
and the committed tests/golden/sampler_*.npz).

Two executions of the same algorithm are provided:

* ``*_loop``: the reference's own control flow -- one ``randint`` call and one Python
  iteration per point -- used as the ``cpu_baseline`` "port" so the timed work is what the
  reference really does on a host core;
* vectorised functions: the same draws in the same order (NumPy's legacy ``randint`` with
  ``size=n`` consumes the MT19937 stream element by element exactly like n scalar calls),
  used as the fast checker in parity tests.

RNG model (sampling.py:113, ``np.random.randint(M)``): legacy ``RandomState`` ->
``_rand_int64`` -> masked rejection on 32-bit MT19937 words: ``mask = 2^ceil(log2(M)) - 1``
(smallest all-ones mask >= M-1); draw ``w & mask`` until ``<= M-1``.  ``M == 1`` consumes
no word.  ``masked_rejection`` restates that on a raw word stream so the CUDA "MT stream"
mode can be checked word for word.

Tie rule.  ``np.argsort(g)[::-1]`` (sampling.py:121) is a reversed ascending sort.  Under the
reference's pinned numpy~=1.19.5 (requirements.txt:2) lists of <= 16 elements are
insertion-sorted, i.e. stable, so equal depths come out *later draw first*.  NumPy >= 1.25
uses an unstable SIMD sort, so the reference itself is not reproducible on ties there.  The
oracle (and the CUDA path) fix the rule "depth descending, ties: later draw first"
(= reversed stable argsort) for every K.

Score arithmetic is NumPy-version sensitive (sampling.py:161-169, 194-208, 227-239 mix
np.float32 scalars with Python numbers).  ``promotion='nep50'`` restates what NumPy >= 2
does (float32 throughout; this is what the reference does in this container and what the
goldens pin); ``promotion='legacy'`` restates value-based casting of NumPy 1.x (float64
accumulators) and is unpinned.
"""
import numpy as np

DEFAULT_FACTORS = {"purely": 0.8, "masked": 1.5, "thresholded": 1.5, "information": 5}


# ----------------------------------------------------------------------------- geometry
def squeeze_gt(gt):
    """gt arrives as (H, W) or (H, W, 1) (hourglass_provider.py:43-47, sampling.py:118)."""
    gt = np.asarray(gt)
    if gt.ndim == 3 and gt.shape[-1] == 1:
        gt = gt[..., 0]
    if gt.ndim != 2:
        raise ValueError("gt must be (H, W) or (H, W, 1)")
    return gt


def num_candidates(batch_size, batch_size_factor):
    """sampling.py:55 -- ``int(batch_size * batch_size_factor)``."""
    return int(batch_size * batch_size_factor)


def valid_flat_indices(mask, image_shape):
    """Flat image index of every valid mask pixel, in ``np.where`` (row-major) order.

    sampling.py:135 ``np.where(mask > 0)``; sampling.py:124-129 scales
    ``x = H / Hm``, ``y = W / Wm`` (Python floats); sampling.py:115-119
    ``r = int(rows[sel] * x)``, ``c = int(cols[sel] * y)``, ``p = r * W + c``.
    """
    mask = np.asarray(mask)
    if mask.ndim == 3 and mask.shape[-1] == 1:
        mask = mask[..., 0]
    H, W = int(image_shape[0]), int(image_shape[1])
    rows, cols = np.where(mask > 0)
    x_scale = H / mask.shape[0]
    y_scale = W / mask.shape[1]
    r = (rows * x_scale).astype(np.int64)      # int() truncation of a non-negative float64
    c = (cols * y_scale).astype(np.int64)
    return r * W + c


# ----------------------------------------------------------------------------- RNG model
def rejection_mask(M):
    """Smallest 2^k - 1 >= M - 1 (numpy legacy ``_gen_mask``)."""
    return (1 << int(M - 1).bit_length()) - 1


def masked_rejection(raw_words, M, count):
    """Consume ``raw_words`` (uint32 MT19937 outputs) the way ``randint(M)`` does.

    Returns (selection[count] int64, words_consumed).  Raises if the stream is too short.
    """
    raw_words = np.asarray(raw_words, dtype=np.uint32)
    if M <= 0:
        raise ValueError("randint(0) is an error in the reference too (empty mask)")
    if M == 1:
        return np.zeros(count, np.int64), 0
    v = raw_words & np.uint32(rejection_mask(M))
    ok = v <= np.uint32(M - 1)
    csum = np.cumsum(ok)
    if count == 0:
        return np.zeros(0, np.int64), 0
    if csum.size == 0 or csum[-1] < count:
        raise ValueError("raw stream exhausted: need %d accepted, have %d" % (count, csum[-1] if csum.size else 0))
    consumed = int(np.searchsorted(csum, count)) + 1
    return v[:consumed][ok[:consumed]].astype(np.int64), consumed


def raw_words_from_state(rng, n):
    """n raw 32-bit MT19937 outputs from a legacy RandomState (advances it by n words)."""
    return rng.randint(0, 2 ** 32, size=n, dtype=np.uint32)


def draw_selection(M, count, rng=None):
    """``count`` sequential ``randint(M)`` draws (sampling.py:113) from ``rng`` (default: global)."""
    rng = np.random if rng is None else rng
    return np.asarray(rng.randint(M, size=count), dtype=np.int64)


# ----------------------------------------------------------------------------- core sampler
def rankings_from_selection(sel, valid_flat, gt, K):
    """Build ``(n, K, 2) float32`` rankings from fed valid-pixel selections.

    sampling.py:110-122 + 137-143: per list gather gt at the K drawn pixels, order by gt
    descending (tie rule in the module docstring), store ``(flat index, depth)`` as float32.
    """
    gt = squeeze_gt(gt)
    sel = np.asarray(sel, dtype=np.int64).reshape(-1, K)
    p = np.asarray(valid_flat, dtype=np.int64)[sel]
    g = gt.reshape(-1)[p].astype(np.float64)                     # gts_buffer is float64 (sampling.py:57)
    order = np.argsort(g, axis=1, kind="stable")[:, ::-1]
    out = np.empty(sel.shape + (2,), dtype=np.float32)
    out[..., 0] = np.take_along_axis(p, order, axis=1)
    out[..., 1] = np.take_along_axis(g, order, axis=1)
    return out


def sample_masked_rankings(image_shape, mask, gt, batch_size, batch_size_factor, K, rng=None):
    """sampling.py:131-145 vectorised.  Returns (result (n,K,2) f32, sel (n*K,) int64)."""
    n = num_candidates(batch_size, batch_size_factor)
    valid_flat = valid_flat_indices(mask, image_shape)
    sel = draw_selection(valid_flat.shape[0], n * K, rng)
    return rankings_from_selection(sel, valid_flat, gt, K), sel


def sample_masked_rankings_loop(image_shape, mask, gt, batch_size, batch_size_factor, K, rng=None):
    """sampling.py:110-145 with the reference's own control flow (one draw per iteration).

    This is the "port" timed as ``cpu_baseline``: the same per-point interpreter work as the
    reference (randint call, two scaled int conversions, gt read, index arithmetic, per-list
    argsort), written independently.
    """
    rng = np.random if rng is None else rng
    gt = squeeze_gt(gt)
    mask = np.asarray(mask)
    H, W = int(image_shape[0]), int(image_shape[1])
    x_scale = H / mask.shape[0]
    y_scale = W / mask.shape[1]
    n = num_candidates(batch_size, batch_size_factor)
    out = np.zeros((n, K, 2), dtype=np.float32)
    depth_buf = np.zeros(K)
    index_buf = np.zeros(K)
    rows, cols = np.where(mask > 0)
    M = rows.shape[0]
    for i in range(n):
        for j in range(K):
            s = rng.randint(M)
            r = int(rows[s] * x_scale)
            c = int(cols[s] * y_scale)
            depth_buf[j] = gt[r, c]
            index_buf[j] = r * W + c
        order = np.argsort(depth_buf, kind="stable")[::-1]
        out[i, :, 0] = index_buf[order]
        out[i, :, 1] = depth_buf[order]
    return out


# ----------------------------------------------------------------------------- scores
def _relation_is_equal(g1, g2, threshold, promotion):
    """depth_utils.py:5-21 with a threshold: 0 ("equal") iff the ratio lies strictly inside
    (1/(1+t), 1+t).  g1, g2: float32 arrays."""
    if promotion == "nep50":
        eps = np.float32(1e-10)
        ratio = (g1 + eps) / (g2 + eps)                      # float32 / float32
        hi = np.float32(1 + threshold)                        # weak Python float -> float32
        lo = np.float32(1 / (1 + threshold))
    elif promotion == "legacy":
        ratio = (g1.astype(np.float64) + 1e-10) / (g2.astype(np.float64) + 1e-10)
        hi = 1 + threshold
        lo = 1 / (1 + threshold)
    else:
        raise ValueError(promotion)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ~(ratio >= hi) & ~(ratio <= lo)


def score_adjacent_differences(result, threshold=None, equality_penalty=-1000, promotion="nep50"):
    """sampling.py:161-167 (threshold None) and sampling.py:194-205 (thresholded).

    ``tmp = 0; for j: [tmp += penalty if equal]; tmp += |g_j - g_{j+1}|`` -- sequential,
    float32 (nep50) or float64 (legacy) accumulator; stored into a float64 ``dists`` array.
    """
    g = np.asarray(result)[:, :, 1]
    acc_t = np.float32 if promotion == "nep50" else np.float64
    acc = np.zeros(g.shape[0], dtype=acc_t)
    pen = acc_t(equality_penalty)
    for j in range(g.shape[1] - 1):
        diff = np.abs(g[:, j] - g[:, j + 1])                  # float32
        if threshold is not None:
            eq = _relation_is_equal(g[:, j], g[:, j + 1], threshold, promotion)
            acc = np.where(eq, acc + pen, acc)
        acc = acc + diff.astype(acc_t)
    return acc.astype(np.float64)


def expected_depth_ladder(gt, K, promotion="nep50"):
    """sampling.py:219-223 ``np.linspace(min(gt) + 0.001, max(gt), K + 1)[1:]``."""
    gt = np.asarray(gt)
    lo = np.amin(gt)
    hi = np.amax(gt)
    if promotion == "legacy":
        lo = np.float64(lo)
        hi = np.float64(hi)
    return np.linspace(lo + 0.001, hi, K + 1)[1:]


def score_information(result, gt, threshold=0.03, equality_penalty=-1000, promotion="nep50"):
    """sampling.py:227-237: ``-(sum((g - e)^2 / e))`` (NumPy pairwise sum) then, in the
    float64 ``score_Id`` slot, ``+= penalty`` per "equal" adjacent pair."""
    res = np.asarray(result)
    K = res.shape[1]
    e = expected_depth_ladder(gt, K, promotion)
    g = res[:, :, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        chi = (np.square(g - e) / e)
        score = np.empty(g.shape[0], dtype=np.float64)
        for i in range(g.shape[0]):                           # row-wise .sum() keeps NumPy's own
            score[i] = -(chi[i].sum())                        # pairwise association per list
    for j in range(K - 1):
        eq = _relation_is_equal(g[:, j], g[:, j + 1], threshold, promotion)
        score = np.where(eq, score + equality_penalty, score)
    return score


def select_top(result, scores, batch_size):
    """sampling.py:169/208/239 ``result[np.argsort(scores)[::-1]][:batch_size]``.
    Ties: reversed stable argsort (larger candidate index first), as for the per-list sort."""
    order = np.argsort(np.asarray(scores, dtype=np.float64), kind="stable")[::-1]
    return np.asarray(result)[order][:batch_size], order[:batch_size]


# ----------------------------------------------------------------------------- strategies
def sample_masked_point_batch(strategy, image_shape, mask, gt, batch_size, K, batch_size_factor=None,
                              threshold=0.03, equality_penalty=-1000, rng=None, promotion="nep50",
                              sel=None):
    """The four masked strategies' ``sample_masked_point_batch`` (sampling.py:147-150,
    157-169, 190-208, 218-239).  ``sel`` feeds pre-drawn selections instead of ``rng``
    (used to check the Philox mode of the CUDA path).  Returns (rankings, sel, scores|None).
    """
    f = DEFAULT_FACTORS[strategy] if batch_size_factor is None else batch_size_factor
    if sel is None:
        result, sel = sample_masked_rankings(image_shape, mask, gt, batch_size, f, K, rng)
    else:
        result = rankings_from_selection(sel, valid_flat_indices(mask, image_shape), gt, K)
        assert result.shape[0] == num_candidates(batch_size, f)
    if strategy == "purely":
        return result[:batch_size], sel, None
    if strategy == "masked":
        scores = score_adjacent_differences(result, None, equality_penalty, promotion)
    elif strategy == "thresholded":
        scores = score_adjacent_differences(result, threshold, equality_penalty, promotion)
    elif strategy == "information":
        scores = score_information(result, squeeze_gt(gt), threshold, equality_penalty, promotion)
    else:
        raise ValueError(strategy)
    top, _ = select_top(result, scores, batch_size)
    return top, sel, scores
