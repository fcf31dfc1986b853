#!/usr/bin/env python
"""Platform check for the end-to-end leg: pinned host <-> device copy rates per GPU when 1..N GPUs copy at once
(one thread per GPU).  Two modes per GPU count: H2D only (77 MB per copy like the inputs of one config-2 step) and the
step's own traffic pattern -- 77 MB in and 26 MB out concurrently on two streams (PCIe is full duplex, the host memory
system is not free).  Tells whether the e2e numbers at N > 1 are limited by the host / PCIe fabric rather than by this
library: the bidirectional rows are the ceiling of `e2e` at that GPU count (ms per step >= 77 MB / h2d rate of the
slowest GPU)."""
import json
import threading
import time

import torch

n = torch.cuda.device_count()
MB_IN, MB_OUT = 77, 26
bufs = [(torch.empty(MB_IN << 20, dtype=torch.uint8).pin_memory(),
         torch.empty(MB_IN << 20, dtype=torch.uint8, device="cuda:%d" % i),
         torch.empty(MB_OUT << 20, dtype=torch.uint8).pin_memory(),
         torch.empty(MB_OUT << 20, dtype=torch.uint8, device="cuda:%d" % i)) for i in range(n)]


def worker(i, reps, out, both):
    torch.cuda.set_device(i)
    h_in, d_in, h_out, d_out = bufs[i]
    s_in, s_out = torch.cuda.Stream(i), torch.cuda.Stream(i)

    def burst(k):
        for _ in range(k):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    burst(3)
    t0 = time.perf_counter()
    burst(reps)
    dt = time.perf_counter() - t0
    out[i] = (reps * (MB_IN << 20) / dt / 1e9, reps * (MB_OUT << 20) / dt / 1e9 if both else 0.0, 1e3 * dt / reps)


for both in (False, True):
    for k in sorted(set([1, 2, 4, n])):
        if k > n:
            continue
        out = {}
        th = [threading.Thread(target=worker, args=(i, 40, out, both)) for i in range(k)]
        [t.start() for t in th]
        [t.join() for t in th]
        row = {"gpus_copying": k, "mode": "h2d 77 MB + d2h 26 MB concurrently" if both else "h2d 77 MB",
               "h2d_GBps_per_gpu": [round(out[i][0], 1) for i in range(k)],
               "aggregate_h2d_GBps": round(sum(v[0] for v in out.values()), 1)}
        if both:
            row["d2h_GBps_per_gpu"] = [round(out[i][1], 1) for i in range(k)]
            row["ms_per_step_slowest_gpu"] = round(max(v[2] for v in out.values()), 3)
        print(json.dumps(row))
