"""Evaluation list generators of the reference on the GPU, bit-compatible with the NumPy global stream.

Mirror of pldepth/data/providers/generic_ranking_provider.py:

    prov = GenericHourglassPairRelationDataProvider(model_params, seed, invert_relation_sign, threshold=0.03)
    np.random.seed(seed); pairs = prov.generate_ordinal_pairs(base_ds)        # np.float32 [N, n_pairs, 5]   (80-111)

    prov = GenericHourglassRankingDataProvider(model_params, query_ranking_size, seed, invert_relation_sign)
    np.random.seed(seed); lists = prov.generate_rankings(base_ds)             # np.float32 [N, 100, K, 2]    (180-215)

``base_ds`` is any iterable of ``(image, gt)`` elements (NumPy arrays or tensors; what the reference gets from
``tf.data.Dataset.as_numpy_iterator()``), or an object with ``as_numpy_iterator``.  Consecutive elements of equal
size are processed as one device batch; the MT19937 word cursor is carried from batch to batch on the device.

Random-number modes (``rng``): ``"numpy"`` consumes the global ``np.random`` state exactly as the reference's
``np.random.randint`` calls do and leaves it where the reference would (one synchronisation per call);
``"mt19937"`` generates the same stream on the device from ``seed`` (``np.random.seed(seed)`` equivalent).
The TF dataset plumbing (``Dataset.zip`` / ``cache``) stays with the framework: ``provide_val_dataset`` /
``provide_test_dataset`` seed the stream like the reference (120-121 / 139-140) and return ``(base_ds, lists)``.
"""
import os

import numpy as np
import torch

from . import ops
from ._lib import PROMOTION, check
from .sampling import _DeviceMT19937, _NP_RNG_LOCK, _device


def _elements(base_ds):
    it = base_ds.as_numpy_iterator() if hasattr(base_ds, "as_numpy_iterator") else iter(base_ds)
    for elem in it:
        image, gt = elem[0], elem[1]
        if isinstance(gt, torch.Tensor):
            gt = gt.detach().cpu().numpy()
        gt = np.squeeze(np.asarray(gt, dtype=np.float32))
        if gt.ndim != 2:
            raise ValueError("every ground-truth map must squeeze to [H, W], got %s" % (gt.shape,))
        shape = tuple(int(d) for d in np.shape(image) if int(d) != 1)      # np.squeeze(elem[0]).shape
        yield shape, gt


def _runs(base_ds, need_image_shape):
    """Consecutive elements of equal size -> one float32 [N,H,W] block each (order preserved)."""
    block, size = [], None
    for shape, gt in _elements(base_ds):
        if need_image_shape:
            if shape[:2] != gt.shape:
                # generate_ordinal_pairs draws inside image.shape and indexes gt (generic_ranking_provider.py:91-99)
                raise ValueError("image %s and gt %s must have the same height / width" % (shape, gt.shape))
        if size is not None and gt.shape != size:
            yield np.stack(block)
            block = []
        size = gt.shape
        block.append(gt)
    if block:
        yield np.stack(block)


class _StreamUser(object):
    """Shared word-stream handling of the two providers."""

    def _init_stream(self, seed, rng, device, promotion):
        if rng not in ("numpy", "mt19937"):
            raise ValueError("rng must be 'numpy' or 'mt19937' (these generators exist to reproduce the NumPy stream)")
        self.rng, self.promotion, self._device_arg, self._mt = rng, promotion, device, None
        self.seed = seed

    def reseed(self):
        """np.random.seed(self.seed) (generic_ranking_provider.py:29,47,139,157) for the chosen stream."""
        if self.rng == "numpy":
            np.random.seed(self.seed)
        else:
            self._mt = None

    def _with_words(self, dev, need, fn):
        """Run ``fn(raw i32 tensor, consumed i64[1])`` on a window of ``need``-many accepted draws' worth of raw
        words and account for what it consumed."""
        window = int(need * 2.5) + 8192
        consumed = torch.zeros(1, dtype=torch.int64, device=dev)
        if self.rng == "numpy":
            with _NP_RNG_LOCK:
                st0 = np.random.get_state()
                raw_h = np.random.randint(0, 2 ** 32, size=window, dtype=np.uint32)
                raw = torch.from_numpy(raw_h.view(np.int32)).to(dev)
                try:
                    out = fn(raw, consumed)
                    ops.check_status(dev)
                finally:
                    used = int(consumed.item())
                    np.random.set_state(st0)
                    if used > 0:
                        np.random.randint(0, 2 ** 32, size=min(used, window), dtype=np.uint32)
            return out
        if self._mt is None:
            self._mt = _DeviceMT19937(int(self.seed), dev)
        raw = self._mt.window(window)
        out = fn(raw, consumed)
        ops.check_status(dev)
        self._mt.consume(int(consumed.item()))
        return out


class GenericHourglassPairRelationDataProvider(_StreamUser):
    """generic_ranking_provider.py:12-111 (constructor arguments identical; ``rng`` / ``device`` / ``promotion`` are
    extensions).  ``threshold=None`` compares the depths directly (depth_utils.py:6-12)."""

    def __init__(self, model_params, seed, invert_relation_sign, threshold=0.03, cache_val_data=True,
                 save_pairs_on_disk=False, config=None, rng="numpy", device=None, promotion="nep50"):
        self.model_params = model_params
        self.invert_relation_sign = invert_relation_sign
        self.threshold = threshold
        self.cache_val_data = cache_val_data
        self.dataset_name = model_params.get_parameter("dataset")
        self.save_pairs_on_disk = save_pairs_on_disk
        if self.save_pairs_on_disk:
            assert config is not None, "If the generated pairs should be saved, a configuration specifying the " \
                                       "cache location must be given!"
        self.config = config
        self._init_stream(seed, rng, device, promotion)

    def provide_train_dataset(self, base_ds, base_ds_gts=None):
        raise NotImplementedError("Training provision is not implemented yet.")

    def _cache_path(self, *parts):
        return os.path.join(self.config["DATA"]["CACHE_PATH_PREFIX"],
                            "ordinal_pair_cache/{}.npy".format("_".join(str(p) for p in parts)))

    def provide_val_dataset(self, base_ds, base_ds_gts=None):
        self.reseed()
        path = self._cache_path(self.dataset_name, "val", self.model_params.get_parameter("val_rankings_per_img"),
                                self.seed) if self.save_pairs_on_disk else None
        return base_ds, self.retrieve_ordinal_pairs(base_ds, path)

    def provide_test_dataset(self, base_ds):
        self.reseed()
        path = self._cache_path(self.dataset_name, self.model_params.get_parameter("val_rankings_per_img"),
                                self.seed) if self.save_pairs_on_disk else None
        return base_ds, self.retrieve_ordinal_pairs(base_ds, path)

    def retrieve_ordinal_pairs(self, base_ds, cache_path):
        if not self.save_pairs_on_disk:
            return self.generate_ordinal_pairs(base_ds, invert_relation_sign=self.invert_relation_sign)
        if not os.path.exists(cache_path):
            pairs = self.generate_ordinal_pairs(base_ds, invert_relation_sign=self.invert_relation_sign)
            np.save(cache_path, pairs)
            return pairs
        return np.load(cache_path)

    def generate_ordinal_pairs_device(self, gts, invert_relation_sign=False):
        """gts: float32 [N,H,W] on the device -> float32 [N, val_rankings_per_img, 5] on the device."""
        n_pairs = int(self.model_params.get_parameter("val_rankings_per_img"))
        gts = ops.as_cuda(gts, torch.float32, "gts")
        N, H, W = gts.shape
        dev = gts.device
        ctx = ops._ctx(gts)
        thr = -1.0 if self.threshold is None else float(self.threshold)

        def call(raw, consumed):
            with torch.cuda.device(dev):
                out = torch.empty((N, n_pairs, 5), dtype=torch.float32, device=dev)
                check(ctx.lib.pld_eval_ordinal_pairs_mt(ctx.handle, ops._p(gts), N, H, W, n_pairs, thr,
                                                        1 if invert_relation_sign else 0, PROMOTION[self.promotion],
                                                        ops._p(raw), int(raw.numel()), ops._p(consumed), ops._p(out),
                                                        ops._stream(dev)))
            return out

        if N * n_pairs == 0:
            return torch.zeros((N, n_pairs, 5), dtype=torch.float32, device=dev)
        return self._with_words(dev, 4 * N * n_pairs, call)

    def generate_ordinal_pairs(self, base_ds_imgs_gts, invert_relation_sign=False):
        dev = _device(self._device_arg)
        n_pairs = int(self.model_params.get_parameter("val_rankings_per_img"))
        parts = [self.generate_ordinal_pairs_device(torch.from_numpy(block).to(dev), invert_relation_sign).cpu().numpy()
                 for block in _runs(base_ds_imgs_gts, need_image_shape=True)]
        return np.concatenate(parts) if parts else np.zeros([0, n_pairs, 5], np.float32)


class GenericHourglassRankingDataProvider(_StreamUser):
    """generic_ranking_provider.py:114-215 (constructor arguments identical; ``rng`` / ``device`` are extensions)."""

    def __init__(self, model_params, query_ranking_size, seed, invert_relation_sign, threshold=0.03,
                 cache_val_data=True, save_rankings_on_disk=False, config=None, rng="numpy", device=None):
        self.model_params = model_params
        self.query_ranking_size = query_ranking_size
        self.invert_relation_sign = invert_relation_sign
        self.threshold = threshold
        self.cache_val_data = cache_val_data
        self.dataset_name = model_params.get_parameter("dataset")
        self.save_rankings_on_disk = save_rankings_on_disk
        if self.save_rankings_on_disk:
            assert config is not None, "If the generated rankings should be saved, a configuration specifying the " \
                                       "cache location must be given!"
        self.config = config
        self._init_stream(seed, rng, device, "nep50")

    def provide_train_dataset(self, base_ds, base_ds_gts=None):
        raise NotImplementedError("Providing training data is not supported.")

    def _cache_path(self, *parts):
        return os.path.join(self.config["DATA"]["CACHE_PATH_PREFIX"],
                            "ranking_cache/{}.npy".format("_".join(str(p) for p in parts)))

    def provide_val_dataset(self, base_ds, base_ds_gts=None):
        self.reseed()
        path = self._cache_path(self.dataset_name, "val", 100, self.seed, self.query_ranking_size) \
            if self.save_rankings_on_disk else None
        return base_ds, self.retrieve_rankings(base_ds, path)

    def provide_test_dataset(self, base_ds):
        self.reseed()
        path = self._cache_path(self.dataset_name, 100, self.seed, self.query_ranking_size) \
            if self.save_rankings_on_disk else None
        return base_ds, self.retrieve_rankings(base_ds, path)

    def retrieve_rankings(self, base_ds, cache_path):
        if not self.save_rankings_on_disk:
            return self.generate_rankings(base_ds, invert_relation_sign=self.invert_relation_sign)
        if not os.path.exists(cache_path):
            rankings = self.generate_rankings(base_ds, invert_relation_sign=self.invert_relation_sign)
            np.save(cache_path, rankings)
            return rankings
        return np.load(cache_path)

    def generate_rankings_device(self, gts, invert_relation_sign=False, val_rankings_per_img=100):
        """gts: float32 [N,H,W] on the device -> float32 [N, val_rankings_per_img, K, 2] on the device."""
        K, n = int(self.query_ranking_size), int(val_rankings_per_img)
        gts = ops.as_cuda(gts, torch.float32, "gts")
        N, HW = gts.shape[0], gts.shape[1] * gts.shape[2]
        dev = gts.device
        if N * n == 0:
            return torch.zeros((N, n, K, 2), dtype=torch.float32, device=dev)
        # np.random.randint(0, len(gt)) over the flattened map = the core sampler on the identity table
        n_valid = torch.full((N,), -HW, dtype=torch.int32, device=dev)
        unused = torch.empty((N, 1), dtype=torch.int32, device=dev)

        def call(raw, consumed):
            rankings, _ = ops.sample_lists_mt(gts.reshape(N, HW), unused, n_valid, K, n, raw, consumed)
            if invert_relation_sign:
                ctx = ops._ctx(gts)
                with torch.cuda.device(dev):
                    check(ctx.lib.pld_eval_invert_rankings(ctx.handle, ops._p(rankings), N * n, K, ops._stream(dev)))
            return rankings

        if HW == 1:                                   # randint(0, 1) consumes no word: every draw is pixel 0
            z = torch.zeros(1, dtype=torch.int32, device=dev)
            return call(z, torch.zeros(1, dtype=torch.int64, device=dev))
        return self._with_words(dev, N * n * K, call)

    def generate_rankings(self, base_ds_imgs_gts, invert_relation_sign=False, val_rankings_per_img=100):
        dev = _device(self._device_arg)
        parts = [self.generate_rankings_device(torch.from_numpy(block).to(dev), invert_relation_sign,
                                               val_rankings_per_img).cpu().numpy()
                 for block in _runs(base_ds_imgs_gts, need_image_shape=False)]
        if not parts:
            return np.zeros([0, val_rankings_per_img, int(self.query_ranking_size), 2], np.float32)
        return np.concatenate(parts)
