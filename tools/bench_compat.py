#!/usr/bin/env python
"""Timing of the NumPy-compatible paths: device MT19937 generator and the per-image drop-in sampler call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pldepth_b200 import ops, sampling, synth  # noqa: E402
from pldepth_b200.models_meta import ModelParameters  # noqa: E402

dev = torch.device("cuda", 0)
state, pos = ops.mt19937_init(1, dev)
ops.mt19937_generate(state, pos, 1 << 20)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 1 << 26
ops.mt19937_generate(state, pos, n)
torch.cuda.synchronize()
print("device MT19937: %.2f G words/s" % (n / (time.perf_counter() - t0) / 1e9))

H = W = 448
gt = synth.depth_map(H, W, 1)
mask = synth.valid_mask(H, W, 2, 0.1)
image = np.zeros((H, W, 3), np.float32)
for cls, rng in ((sampling.ThresholdedMaskedRandomSamplingStrategy, "numpy"),
                 (sampling.InformationScoreBasedSampling, "numpy"),
                 (sampling.ThresholdedMaskedRandomSamplingStrategy, "philox")):
    s = cls(ModelParameters(ranking_size=5), rng=rng)
    for R in (100, 1000, 10000):
        s.sample_masked_point_batch(image, mask, gt, R)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            s.sample_masked_point_batch(image, mask, gt, R)
        dt = (time.perf_counter() - t0) / reps
        print("%s rng=%s R=%d: %.2f ms per image call (%.3g kept lists/s)" % (cls.__name__, rng, R, dt * 1e3, R / dt))
