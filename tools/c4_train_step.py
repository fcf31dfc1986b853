#!/usr/bin/env python
"""BASELINE config 4: an ff_effnet PLDepth training step (448x448, batch 64, ranking_size 5) with the
sampler + loss replaced by the fused CUDA step, per-image sharding over N GPUs (torchrun / DDP).

The encoder-decoder is the framework's business (north_star): torchvision EfficientNet-B0 (random
init, no weights offline) + the 5-stage conv/BN/ReLU/bilinear decoder with the three expand-activation
skips of pldepth/models/pl_hourglass.py:43-100, run by PyTorch/cuDNN in bf16 autocast.  What this
script shows is the share of the step spent in the PL path: it prints one JSON line with the full
training-step time, the network-only time (same step fed a constant gradient) and the PL-path time.

  python tools/c4_train_step.py [--batch 64] [--rankings 1000] [--steps 10]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c4_train_step.py --batch 64
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class EffNetHourglass(nn.Module):
    def __init__(self):
        super().__init__()
        import torchvision
        self.features = torchvision.models.efficientnet_b0(weights=None).features
        self._taps = {}
        # expand activations of the first block of stages 3, 4 and 6 (Keras block3a/4a/6a_expand_activation)
        for name, idx in (("s3", 3), ("s4", 4), ("s6", 6)):
            self.features[idx][0].block[0].register_forward_hook(self._make_hook(name))

        def stage(cin, cout):
            return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
        self.d0 = stage(1280, 672)
        self.d1 = stage(672 + 672, 240)
        self.d2 = stage(240 + 240, 144)
        self.d3 = stage(144 + 144, 32)
        self.d4 = stage(32, 32)
        self.head = nn.Conv2d(32, 1, 3, padding=1)

    def _make_hook(self, name):
        def hook(_m, _i, out):
            self._taps[name] = out
        return hook

    def forward(self, x):
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
        e = self.features(x)
        x = up(self.d0(e))
        x = up(self.d1(torch.cat([x, self._taps["s6"]], 1)))
        x = up(self.d2(torch.cat([x, self._taps["s4"]], 1)))
        x = up(self.d3(torch.cat([x, self._taps["s3"]], 1)))
        x = up(self.d4(x))
        return self.head(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--rankings", type=int, default=1000)
    ap.add_argument("--ranking-size", type=int, default=5)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--strategy", default="purely", choices=["purely", "masked", "thresholded", "information"])
    args = ap.parse_args()
    import numpy as np
    import torch.distributed as dist
    from pldepth_b200 import synth
    from pldepth_b200.dist import shard_bounds
    from pldepth_b200.losses import SampledHourglassNLL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(args.batch, rank, world)
    B, H, W, K, R = hi - lo, args.size, args.size, args.ranking_size, args.rankings
    torch.manual_seed(0)
    model = EffNetHourglass().to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, amsgrad=True)      # PLDepth.py:133
    base = synth.depth_map(H, W, 4000 + rank)
    gt = torch.from_numpy(np.stack([np.roll(base, 13 * b, axis=1) for b in range(B)])).to(dev)
    mask = torch.ones((B, H, W), dtype=torch.float32, device=dev)
    images = torch.randn((B, 3, H, W), device=dev).to(memory_format=torch.channels_last)
    criterion = SampledHourglassNLL(K, R, strategy=args.strategy, seed=4, global_batch=args.batch, image_base=lo)
    pl = criterion.step

    def train_step(with_pl):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = model(images)
        pred = pred.float().contiguous()                 # (B,1,H,W) == (B,H,W,1) in memory
        if with_pl:
            loss = criterion(gt, mask, pred)             # sampler + gather + PL loss; backward = its dense gradient
            loss.backward()
            losses.append(loss.detach())
        else:
            pred.backward(gradient=torch.full_like(pred, 1e-6))
        opt.step()

    losses = []

    def timed(fn, n):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    t_full = timed(lambda: train_step(True), args.steps)
    t_net = timed(lambda: train_step(False), args.steps)
    pred0 = torch.randn((B, H, W, 1), device=dev)
    t_pl = timed(lambda: pl.run(gt, mask, pred0), max(args.steps, 50))
    loss_first = torch.stack(losses[:args.warmup]).mean()
    loss_last = torch.stack(losses[-args.warmup:]).mean()
    if world > 1:                                        # each rank holds its share of the global mean
        both = torch.stack([loss_first, loss_last])
        dist.all_reduce(both)
        loss_first, loss_last = both[0], both[1]
    if rank == 0:
        print(json.dumps({"config": "C4 ff_effnet training step, %dx%d, global batch %d over %d GPU(s), K=%d, R=%d, "
                                    "sampler %s" % (H, W, args.batch, world, K, R, args.strategy),
                          "loss_first_steps": float(loss_first), "loss_last_steps": float(loss_last),
                          "train_step_ms": t_full, "network_only_ms": t_net, "pl_path_ms": t_pl,
                          "pl_share_of_step": t_pl / t_full, "lists_per_step": args.batch * R,
                          "images_per_s": args.batch / (t_full * 1e-3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
