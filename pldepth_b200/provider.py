"""Data-provider side of the hot path (mirror of pldepth/data/providers/hourglass_provider.py).

The reference wraps the per-image sampler in ``tf.numpy_function`` inside a tf.data map, then
``.batch(B, drop_remainder=True).prefetch().repeat()`` (hourglass_provider.py:29-62), and
pre-generates fixed validation rankings with the thresholded sampler (64-73, 179-193).

Here:
  * ``sample_rankings`` keeps the per-image NumPy contract (hourglass_provider.py:75-86) so the class
    can be dropped behind ``tf.numpy_function`` unchanged;
  * ``sample_rankings_batch`` samples a whole batch on the device in one go (what the map + batch
    produce: ``y_true (B, R, K, 2) float32``) -- no per-image Python callback, no GIL;
  * ``iterate_train_batches`` is the tf.data pipeline for in-memory arrays: shuffle, optional
    consistent left-right flip of image / mask / gt (34-51), batches with drop_remainder, repeat;
  * ``generate_validation_rankings`` produces the fixed ``[N, val_R, K, 2]`` array (179-193).
"""
import numpy as np
import torch

from .sampling import ThresholdedMaskedRandomSamplingStrategy


class HourglassLargeScaleDataProvider(object):
    def __init__(self, model_params, train_consistency_masks=None, val_consistency_masks=None, loss_type="NLL",
                 augmentation=False, sampling_eq_threshold=0.03, bs_factor=5, rng="numpy", seed=0):
        self.model_params = model_params
        self.train_consistency_masks = train_consistency_masks
        self.val_consistency_masks = val_consistency_masks
        # hourglass_provider.py:21-22: both default samplers are the thresholded strategy
        self.random_sampler = ThresholdedMaskedRandomSamplingStrategy(model_params, sampling_eq_threshold, rng=rng,
                                                                      seed=seed)
        self.val_random_sampler = ThresholdedMaskedRandomSamplingStrategy(model_params, rng=rng, seed=seed + 1)
        self.augmentation = augmentation
        self.loss_type = loss_type
        self.bs_factor = bs_factor

    # ---- per-image contract (hourglass_provider.py:75-86) ---------------------------------------
    def sample_rankings(self, image, cons_mask, gt, sampling_strategy=None, rankings_per_img=None,
                        return_image=True):
        if sampling_strategy is None:
            sampling_strategy = self.model_params.get_parameter("sampling_strategy")
        if rankings_per_img is None:
            rankings_per_img = self.model_params.get_parameter("rankings_per_image")
        result = sampling_strategy.sample_masked_point_batch(image, cons_mask, gt, rankings_per_img)
        if not return_image:
            return result.astype(np.float32)
        return np.asarray(image).astype(np.float32), result.astype(np.float32)

    # ---- whole batch on the device ------------------------------------------------------------------
    def sample_rankings_batch(self, gt, cons_mask, sampling_strategy=None, rankings_per_img=None, image_base=0):
        """gt [B,H,W(,1)], cons_mask [B,Hm,Wm] device tensors -> rankings [B, R_out, K, 2] (device)."""
        if sampling_strategy is None:
            sampling_strategy = self.model_params.get_parameter("sampling_strategy")
        if rankings_per_img is None:
            rankings_per_img = self.model_params.get_parameter("rankings_per_image")
        return sampling_strategy.sample_batch(gt, cons_mask, rankings_per_img, image_base=image_base)

    def iterate_train_batches(self, images, masks, gts, device, shuffle_seed=0, repeat=True):
        """Generator of (images [B,H,W,3], rankings [B,R,K,2]) device tensors.

        ``images/masks/gts``: array-likes indexed by sample.  Mirrors provide_train_dataset
        (hourglass_provider.py:29-62): shuffle, per-sample random left-right flip applied to image,
        mask and gt alike when ``augmentation`` is on, sampler, ``batch(B, drop_remainder=True)``,
        ``repeat()``."""
        B = int(self.model_params.get_parameter("batch_size"))
        n = len(images)
        rng = np.random.RandomState(shuffle_seed)
        while True:
            order = rng.permutation(n)
            for s in range(0, n - B + 1, B):
                idx = order[s:s + B]
                img = np.stack([np.asarray(images[i], dtype=np.float32) for i in idx])
                msk = np.stack([np.squeeze(np.asarray(masks[i], dtype=np.float32)) for i in idx])
                gt = np.stack([np.squeeze(np.asarray(gts[i], dtype=np.float32)) for i in idx])
                if self.augmentation:
                    flip = rng.rand(B) > 0.5
                    img[flip] = img[flip][:, :, ::-1]
                    msk[flip] = msk[flip][:, :, ::-1]
                    gt[flip] = gt[flip][:, :, ::-1]
                img_d = torch.from_numpy(np.ascontiguousarray(img)).to(device, non_blocking=True)
                msk_d = torch.from_numpy(np.ascontiguousarray(msk)).to(device, non_blocking=True)
                gt_d = torch.from_numpy(np.ascontiguousarray(gt)).to(device, non_blocking=True)
                yield img_d, self.sample_rankings_batch(gt_d, msk_d)
            if not repeat:
                return

    def generate_validation_rankings(self, samples):
        """``samples``: iterable of (image, mask, gt) NumPy triples -> float32 [N, val_R, K, 2]
        (hourglass_provider.py:179-193, thresholded sampler, fixed once)."""
        val_r = self.model_params.get_parameter("val_rankings_per_img")
        K = self.model_params.get_parameter("ranking_size")
        samples = list(samples)
        result = np.zeros([len(samples), val_r, K, 2], np.float32)
        for i, (image, mask, gt) in enumerate(samples):
            result[i] = self.sample_rankings(image, mask, gt, self.val_random_sampler, val_r, return_image=False)
        return result
