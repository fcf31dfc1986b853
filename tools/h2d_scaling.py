#!/usr/bin/env python
"""Platform check for the end-to-end leg: pinned host -> device copy rate per GPU when 1..N GPUs copy at once
(one thread per GPU, 77 MB per copy like one config-2 step).  Tells whether the e2e numbers at N > 1 are limited by
the host / PCIe fabric rather than by this library."""
import json
import sys
import threading
import time

import torch

n = torch.cuda.device_count()
MB = 77
bufs = [(torch.empty(MB << 20, dtype=torch.uint8).pin_memory(), torch.empty(MB << 20, dtype=torch.uint8, device="cuda:%d" % i))
        for i in range(n)]


def worker(i, reps, out):
    torch.cuda.set_device(i)
    h, d = bufs[i]
    s = torch.cuda.Stream(i)
    with torch.cuda.stream(s):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        s.synchronize()
    out[i] = reps * (MB << 20) / (time.perf_counter() - t0) / 1e9


for k in sorted(set([1, 2, 4, n])):
    if k > n:
        continue
    out = {}
    th = [threading.Thread(target=worker, args=(i, 40, out)) for i in range(k)]
    [t.start() for t in th]
    [t.join() for t in th]
    print(json.dumps({"gpus_copying": k, "h2d_GBps_per_gpu": [round(out[i], 1) for i in range(k)],
                      "aggregate_GBps": round(sum(out.values()), 1)}))
