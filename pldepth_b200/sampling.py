"""Ranking samplers on the GPU with the reference's interface.

Mirror of pldepth/data/sampling.py (class names, constructor arguments, method names,
argument meaning, return shapes/dtypes and error behaviour of the *masked* strategies that
are on the training path, sampling.py:106-243):

    sampler = ThresholdedMaskedRandomSamplingStrategy(model_params)       # sampling.py:179
    rankings = sampler.sample_masked_point_batch(image, mask, gt, R)      # np.float32 [R, K, 2]

``rankings[:, :, 0]`` = flat pixel index ``r * W + c``, ``[:, :, 1]`` = ground-truth depth,
each list depth-descending; score-based strategies return the R best of ``int(R * factor)``
candidates ordered by score descending.  ``PurelyMasked...`` keeps the reference's quirk of
returning only ``int(0.8 * R)`` lists (sampling.py:147-150).

Random-number modes (constructor keyword ``rng``, an extension):
  * ``"numpy"`` (default): consumes the *global* ``np.random`` MT19937 state exactly as the
    reference's ``np.random.randint(M)`` calls do (sampling.py:113) -- same seed, same lists,
    same state afterwards.  Bit-exact drop-in on tie-free depth maps; synchronises once per call.  (Ties -- equal
    depths inside a list, equal candidate scores -- follow the reversed STABLE argsort: later draw / larger candidate
    index first.  The reference's ``np.argsort(...)[::-1]`` is unstable for more than 16 elements, so on quantised
    depth maps its tie order is build-dependent and may differ.)
  * ``"mt19937"``: same stream, generated on the device from ``seed`` (``np.random.seed(seed)``
    equivalent) -- no host RNG involved.
  * ``"philox"``: counter-based Philox4x32-10 keyed by ``seed``; the throughput mode
    (no synchronisation, used by ``sample_batch`` / the fused training step).

The extra method ``sample_batch(gt, mask, R)`` takes device tensors ``[B,H,W]`` and returns
device rankings ``[B,R,K,2]`` for a whole batch at once (what the reference's tf.data map +
``.batch(B)`` produces, hourglass_provider.py:55-62).
"""
import threading

import numpy as np
import torch

from . import ops

_NP_RNG_LOCK = threading.Lock()


def _device(device):
    if not torch.cuda.is_available():
        from ._lib import PLDError
        raise PLDError("pldepth_b200 needs a CUDA device (no CPU fallback)")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class _DeviceMT19937(object):
    """np.random.seed(seed)-compatible word stream living on the device."""

    def __init__(self, seed, device):
        self.device = device
        self.state, self.pos = ops.mt19937_init(seed, device)
        self.buf = torch.empty(0, dtype=torch.int32, device=device)

    def window(self, n):
        if self.buf.numel() < n:
            fresh = ops.mt19937_generate(self.state, self.pos, n - self.buf.numel())
            self.buf = torch.cat([self.buf, fresh])
        return self.buf

    def consume(self, n):
        self.buf = self.buf[n:].clone()


class SamplingStrategy(object):
    """sampling.py:7-45."""

    def __init__(self, model_params):
        self.num_points_per_sample = model_params.get_parameter('ranking_size')

    @property
    def num_points_per_sample(self):
        return self._num_points_per_sample

    @num_points_per_sample.setter
    def num_points_per_sample(self, value):
        self._num_points_per_sample = value

    def __str__(self):
        return "{}(num_points_per_sample={})".format(self.__class__.__name__, self._num_points_per_sample)


class RandomSamplingStrategy(SamplingStrategy):
    """sampling.py:48-63.  The unmasked ``sample_points*`` methods (sampling.py:65-103) are not
    on the training path (and use the removed ``np.int``); they are intentionally absent."""

    def __init__(self, model_params):
        super(RandomSamplingStrategy, self).__init__(model_params)
        self.threshold = 0.03
        self.downscaling_factor = model_params.get_parameter("downscaling_factor")

    def sample_points(self, image, gt):
        raise NotImplementedError("unmasked sampling is outside the accelerated path (SURVEY.md §8c)")

    def sample_points_batch(self, image, gt, batch_size, batch_size_factor=1.5):
        raise NotImplementedError("unmasked sampling is outside the accelerated path (SURVEY.md §8c)")


class PurelyMaskedRandomSamplingStrategy(RandomSamplingStrategy):
    """sampling.py:106-150."""

    _strategy = "purely"
    _default_factor = 0.8

    def __init__(self, model_params, rng="numpy", seed=0, device=None, promotion="nep50"):
        super().__init__(model_params)
        if rng not in ("numpy", "mt19937", "philox"):
            raise ValueError("rng must be 'numpy', 'mt19937' or 'philox'")
        self.rng = rng
        self.seed = int(seed)
        self.promotion = promotion
        self._device_arg = device
        self._calls = 0            # Philox offset: one fresh sub-stream per call
        self._mt = None
        self._lock = threading.Lock()

    # ---- reference helpers -------------------------------------------------------------
    @staticmethod
    def determine_x_y_scales(image, mask):
        return image.shape[0] / mask.shape[0], image.shape[1] / mask.shape[1]

    # ---- device-native batched API -------------------------------------------------------
    def draw_candidates(self, gt, mask, n, image_shape=None, image_base=0):
        """gt [B,H,W] / mask [B,Hm,Wm] device tensors -> candidate rankings [B,n,K,2] on the
        device (sample_masked_rankings for every image of the batch)."""
        K = int(self._num_points_per_sample)
        gt = ops.as_cuda(gt, torch.float32, "gt")
        if gt.dim() == 4 and gt.shape[-1] == 1:
            gt = gt[..., 0]
        B, H, W = gt.shape
        if image_shape is not None and (int(image_shape[0]) != H or int(image_shape[1]) != W):
            raise ValueError("gt and image must have the same height/width (gt %dx%d, image %dx%d)"
                             % (H, W, image_shape[0], image_shape[1]))
        valid_flat, n_valid = ops.mask_compact(mask, H, W)
        dev = gt.device
        if self.rng == "philox":
            with self._lock:
                off = self._calls
                self._calls += 1
            rankings, _ = ops.sample_lists_philox(gt, valid_flat, n_valid, K, n, self.seed, off, image_base)
            return rankings
        need = B * n * K
        window = int(need * 2.5) + 8192 * B
        consumed = torch.zeros(1, dtype=torch.int64, device=dev)
        if self.rng == "numpy":
            with _NP_RNG_LOCK:
                st0 = np.random.get_state()
                raw_h = np.random.randint(0, 2 ** 32, size=window, dtype=np.uint32)
                raw = torch.from_numpy(raw_h.view(np.int32)).to(dev)
                rankings, _ = ops.sample_lists_mt(gt, valid_flat, n_valid, K, n, raw, consumed)
                try:
                    ops.check_status(dev)
                finally:
                    used = int(consumed.item())
                    np.random.set_state(st0)
                    if used > 0:
                        np.random.randint(0, 2 ** 32, size=min(used, window), dtype=np.uint32)
            return rankings
        with self._lock:
            if self._mt is None:
                self._mt = _DeviceMT19937(self.seed, dev)
            raw = self._mt.window(window)
            rankings, _ = ops.sample_lists_mt(gt, valid_flat, n_valid, K, n, raw, consumed)
            ops.check_status(dev)
            self._mt.consume(int(consumed.item()))
        return rankings

    def score_candidates(self, rankings, gt):
        return None

    def sample_batch(self, gt, mask, batch_size, batch_size_factor=None, image_shape=None, image_base=0):
        """Whole-batch device sampling: returns rankings [B, n_out, K, 2] (device)."""
        f = self._default_factor if batch_size_factor is None else batch_size_factor
        n = int(batch_size * f)
        K = int(self._num_points_per_sample)
        if self.rng == "philox" and self._strategy != "purely" and n >= 1:
            # fast path: score-only pass, radix top-R, kept lists redrawn from their Philox ids
            gt3 = ops.as_cuda(gt, torch.float32, "gt")
            if gt3.dim() == 4 and gt3.shape[-1] == 1:
                gt3 = gt3[..., 0]
            if image_shape is not None and (int(image_shape[0]) != gt3.shape[1] or int(image_shape[1]) != gt3.shape[2]):
                raise ValueError("gt and image must have the same height/width")
            with self._lock:
                off = self._calls
                self._calls += 1
            thr, pen = getattr(self, "threshold", 0.03), getattr(self, "equality_penalty", -1000)
            out = ops.fused_step_scored(mask, gt3.contiguous(), None, K, n, min(batch_size, n), self._strategy, thr, pen,
                                        self.promotion, self.seed, off, image_base)
            return out["rankings"]
        cand = self.draw_candidates(gt, mask, n, image_shape, image_base)
        scores = self.score_candidates(cand, gt)
        if scores is None:
            return cand[:, :batch_size]
        top, _ = ops.select_top(scores, cand, min(batch_size, n))
        return top

    # ---- reference (per-image, NumPy in / NumPy out) interface ---------------------------
    def _to_device(self, mask, gt):
        dev = _device(self._device_arg)
        gt = np.asarray(gt, dtype=np.float32)
        if gt.ndim == 3 and gt.shape[-1] == 1:
            gt = gt[..., 0]
        mask = np.asarray(mask, dtype=np.float32)
        if mask.ndim == 3 and mask.shape[-1] == 1:
            mask = mask[..., 0]
        return (torch.from_numpy(np.ascontiguousarray(gt)).to(dev)[None],
                torch.from_numpy(np.ascontiguousarray(mask)).to(dev)[None])

    def sample_masked_rankings(self, image, mask, gt, batch_size, batch_size_factor=0.8):
        """sampling.py:131-145: (result float32 [int(R*f), K, 2], dists float64 zeros)."""
        n = int(batch_size * batch_size_factor)
        gt_d, mask_d = self._to_device(mask, gt)
        cand = self.draw_candidates(gt_d, mask_d, n, image.shape)
        ops.check_status(cand.device)
        return cand[0].cpu().numpy(), np.zeros(n)

    def sample_masked_point_batch(self, image, mask, gt, batch_size, batch_size_factor=None):
        gt_d, mask_d = self._to_device(mask, gt)
        out = self.sample_batch(gt_d, mask_d, batch_size, batch_size_factor, image.shape)
        ops.check_status(out.device)
        return out[0].cpu().numpy()


class MaskedRandomSamplingStrategy(PurelyMaskedRandomSamplingStrategy):
    """sampling.py:153-169: score = sum of adjacent depth differences, keep the R best of 1.5R."""

    _strategy = "masked"
    _default_factor = 1.5

    def score_candidates(self, rankings, gt):
        return ops.score_lists(rankings, self._strategy, promotion=self.promotion)


class ThresholdedMaskedRandomSamplingStrategy(MaskedRandomSamplingStrategy):
    """sampling.py:172-208: as above, minus ``equality_penalty`` per "equal" adjacent pair."""

    _strategy = "thresholded"

    def __init__(self, model_params, threshold=0.03, equality_penalty=-1000, **kw):
        super().__init__(model_params, **kw)
        self.threshold = threshold
        self.equality_penalty = equality_penalty

    def score_candidates(self, rankings, gt):
        return ops.score_lists(rankings, self._strategy, self.threshold, self.equality_penalty, self.promotion)


class InformationScoreBasedSampling(MaskedRandomSamplingStrategy):
    """sampling.py:211-243: chi-square score against a uniform depth ladder, best R of 5R."""

    _strategy = "information"
    _default_factor = 5

    def __init__(self, model_params, threshold=0.03, equality_penalty=-1000, **kw):
        super().__init__(model_params, **kw)
        self.threshold = threshold
        self.equality_penalty = equality_penalty

    def score_candidates(self, rankings, gt):
        mm = ops.gt_minmax(gt)
        return ops.score_lists(rankings, self._strategy, self.threshold, self.equality_penalty, self.promotion,
                               minmax=mm)

    def __str__(self):
        return "{}(num_points_per_sample={}, threshold={})".format(self.__class__.__name__,
                                                                   self._num_points_per_sample, self.threshold)
