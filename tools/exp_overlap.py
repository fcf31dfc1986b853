#!/usr/bin/env python
"""Experiment: do two independent fused steps on two streams (two contexts, hence two scratch tables) overlap the
HBM-bound table build of one with the request-rate-bound list kernel of the other?"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pldepth_b200 import synth  # noqa: E402
from pldepth_b200._lib import Context, check, c_void_p  # noqa: E402
from pldepth_b200.step import FusedPLStep  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W, K, R = 32, 448, 448, 5, 100000
base = synth.depth_map(H, W, 7)
gt_h = np.stack([np.roll(base, 31 * b, axis=1) for b in range(B)])
sets = []
for s in range(4):
    sets.append(dict(gt=torch.from_numpy(np.roll(gt_h, s, axis=0)).to(dev), mask=torch.ones((B, H, W), device=dev),
                     pred=torch.randn((B, H, W, 1), device=dev),
                     out=FusedPLStep.new_buffers(B, H, W, H, W, R, K, dev)))
ctxs = [Context(0), Context(0)]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
p = lambda t: c_void_p(t.data_ptr())


def step(i, lane):
    s = sets[i % 4]
    ctx = ctxs[lane]
    o = s["out"]
    check(ctx.lib.pld_fused_step(ctx.handle, p(s["mask"]), p(s["gt"]), p(s["pred"]), B, H, W, H, W, K, R, 1, i, 0,
                                 ctypes.c_float(1.0 / (B * R)), p(o["n_valid"]), p(o["rankings"]), p(o["loss"]),
                                 p(o["loss_sum"]), c_void_p(None), p(o["grad"]), c_void_p(streams[lane].cuda_stream)))


def run(n, lanes):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in streams[:lanes]:
        st.wait_event(e0)
    for i in range(n):
        step(i, i % lanes)
    for st in streams[:lanes]:
        ev = torch.cuda.Event()
        ev.record(st)
        torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for lanes in (1, 2, 1, 2):
    run(8, lanes)
    print("streams=%d  %.4f ms/step  %.3e lists/s" % (lanes, run(60, lanes), B * R / (run(60, lanes) * 1e-3)))
