#!/usr/bin/env python
"""Secondary rows of SURVEY.md §8d: throughput of the score-based strategies (kept lists per second)
for BASELINE config-2 shapes: candidates -> scores -> top-R -> ListMLE fwd+bwd, all on one GPU."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--R", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-emit", action="store_true", help="do not materialise the rankings (fused training step)")
    ap.add_argument("--graph", action="store_true", help="replay each fused step from a CUDA graph (FusedPLStep.capture)")
    args = ap.parse_args()
    from pldepth_b200 import ops, sampling, synth
    from pldepth_b200._lib import Context
    from pldepth_b200.models_meta import ModelParameters
    dev = torch.device("cuda", 0)
    B, H, W, K, R = args.B, args.size, args.size, args.K, args.R
    base = synth.depth_map(H, W, 7)
    gt = torch.from_numpy(np.stack([np.roll(base, 31 * b, axis=1) for b in range(B)])).to(dev)
    mask = torch.ones((B, H, W), dtype=torch.float32, device=dev)
    pred = torch.randn((B, H, W, 1), device=dev)
    mp = ModelParameters(ranking_size=K)
    for name, cls in (("purely(f=1.0)", sampling.PurelyMaskedRandomSamplingStrategy),
                      ("masked", sampling.MaskedRandomSamplingStrategy),
                      ("thresholded", sampling.ThresholdedMaskedRandomSamplingStrategy),
                      ("information", sampling.InformationScoreBasedSampling)):
        s = cls(mp, rng="philox", seed=1)
        f = 1.0 if name.startswith("purely") else None

        def step(s=s, f=f, name=name):
            if name.startswith("purely"):
                y = s.sample_batch(gt, mask, R, f)
                return ops.listmle_fwd_bwd(y, pred, B, K, 1.0 / (B * y.shape[1]))
            n = int(R * s._default_factor)          # one call: score pass, top-R, redraw + loss + gradient
            return ops.fused_step_scored(mask, gt, pred, K, n, R, s._strategy, 0.03, -1000, "nep50", seed=1,
                                         offset=step.i, want_rankings=not args.no_emit)
        step.i = 0
        if args.graph and not name.startswith("purely"):
            from pldepth_b200.step import FusedPLStep
            fs = FusedPLStep(K, R, seed=1, strategy=s._strategy, emit_rankings=not args.no_emit)
            graph, _ = fs.capture(gt, mask, pred)
            step = graph.replay
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        l0 = Context.current(0).lib.pld_launch_count()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b_.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / args.steps
        launches = (Context.current(0).lib.pld_launch_count() - l0) / args.steps
        if args.graph:
            Context.current(0).device_offset(False)
        print(json.dumps({"strategy": name, "cuda_graph": bool(args.graph and not name.startswith("purely")),
                          "rankings": "not materialised" if args.no_emit else "emitted", "candidate_factor": s._default_factor if f is None else f,
                          "kept_lists_per_s": B * R / (ms * 1e-3), "ms_per_step": ms, "launches_per_step": launches,
                          "shape": "B=%d %dx%d K=%d R=%d" % (B, H, W, K, R)}))


if __name__ == "__main__":
    main()
