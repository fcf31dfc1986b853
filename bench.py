#!/usr/bin/env python
"""Benchmark of the PLDepth hot path (ranking sampling -> gather -> ListMLE fwd+bwd).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

One "step" = one pass of the fused path over one batch of synthetic maps: valid-pixel table,
Philox sampling of R lists/image ordered by GT depth, emission of the rankings, gather of the
predictions, Plackett-Luce NLL and the dense gradient (scatter-add).  At N GPUs every rank owns
its own B images (per-image sharding, weak scaling); the only collective is the all-reduce of
the scalar loss sum.  Prints ONE JSON line (rank 0).

--impl reference times the reference's own algorithm on the host CPU cores: the oracle "port"
of pldepth/data/sampling.py (same per-point Python loop as the reference) + the NumPy ListMLE
restatement, one process per core, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ranked lists/sec (sample+PL loss fwd+bwd)"
UNIT = "lists/s"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C5s", "C5"])
    ap.add_argument("--sets", type=int, default=6, help="rotating input/output buffer sets (> L2 in total)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work for cpu_baseline")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph")
    ap.add_argument("--lanes", type=int, default=None,
                    help="independent batches in flight on separate streams (each lane has its own context, hence "
                         "its own lookup tables): the HBM-bound table build of one step overlaps the list kernel "
                         "of the previous one.  1 = strictly sequential steps.  Default: 3 (1 for launch-bound "
                         "workloads of fewer than 100 000 lists per step)")
    ap.add_argument("--no-emit", action="store_true", help="do not materialise the rankings (what a fused training "
                                                           "step needs; the default emits them like the reference)")
    ap.add_argument("--hole", type=float, default=0.0, help="fraction of each mask zeroed (default: all-ones mask)")
    return ap.parse_args()


def workload_shape(name):
    from pldepth_b200 import synth
    if name == "C5s":   # a quarter of one GPU's share of config 5 at 8 GPUs (quick run)
        return dict(B=8, H=1024, W=768, K=10, R=1000000)
    if name == "C5":    # one GPU's share of config 5 at 8 GPUs: 32 of the 256 images
        return dict(B=32, H=1024, W=768, K=10, R=1000000)
    return dict(synth.CONFIGS[name])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(B, HW, L, K):
    """SURVEY.md §8d fused op: gt + mask + pred + grad at 4 B/pixel, 8 B per emitted point, loss."""
    return B * HW * 16 + L * K * 8 + 4


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU every 200 ms."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/ code)
# --------------------------------------------------------------------------------------------
def _cpu_image_job(job):
    """One image: reference-style sampler loop + NumPy ListMLE fwd+bwd.  Returns (lists, secs)."""
    import numpy as np
    from oracle import listmle_oracle as lo
    from oracle import sampler_oracle as so
    from pldepth_b200 import synth
    H, W, K, lists, seed = job
    gt = synth.depth_map(H, W, seed)
    mask = synth.valid_mask(H, W, seed, 0.0)
    pred = synth.prediction(H, W, seed + 1)
    rng = np.random.RandomState(seed)
    t0 = time.perf_counter()
    rank = so.sample_masked_rankings_loop((H, W), mask, gt, lists, 1.0, K, rng)
    t1 = time.perf_counter()
    lo.hourglass_nll(rank[None], pred[None], 1, K, dtype=np.float32)
    t2 = time.perf_counter()
    return lists, t1 - t0, t2 - t1


def cpu_rate_single(shape, seconds):
    """cpu_baseline: one core, bounded sample of the same workload (whole images of R lists until about
    `seconds` of CPU work are done)."""
    H, W, K, R = shape["H"], shape["W"], shape["K"], shape["R"]
    probe = _cpu_image_job((H, W, K, 2000, 123))
    per_list = (probe[1] + probe[2]) / probe[0]
    lists_total = int(max(2000, seconds / per_list))
    per_image = min(R, lists_total)
    n_img = max(1, min(shape["B"], int(round(lists_total / per_image))))
    n = ts = tl = 0.0
    for i in range(n_img):
        a, b_, c = _cpu_image_job((H, W, K, per_image, 124 + i))
        n, ts, tl = n + a, ts + b_, tl + c
    return {"value": n / (ts + tl), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d image(s) %dx%d, K=%d, %d lists each: sampler loop %.2fs + ListMLE fwd+bwd (NumPy fp32) %.2fs; "
                      "the reference's sampler is a GIL-bound Python loop, so 1 core is its real rate per "
                      "tf.data worker" % (n_img, H, W, K, per_image, ts, tl),
            "sampler_lists_per_s": n / ts, "loss_lists_per_s": n / tl}


def run_reference(args):
    """--impl reference: every host core runs the port on its own image (best case for the
    reference: tf.data AUTOTUNE parallelism without the GIL)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = workload_shape(args.workload)
    H, W, K = shape["H"], shape["W"], shape["K"]
    cores = os.cpu_count() or 1
    probe = _cpu_image_job((H, W, K, 1000, 5))
    per_list = (probe[1] + probe[2]) / probe[0]
    total_steps = max(1, args.steps + args.warmup)
    budget_per_step = min(20.0, 150.0 / total_steps)
    lists = int(max(200, min(shape["R"], budget_per_step / per_list)))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(i):
            jobs = [(H, W, K, lists, 1000 * i + c) for c in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_cpu_image_job, jobs)
            return sum(r[0] for r in res), time.perf_counter() - t0
        for i in range(args.warmup):
            step(i)
        tot_lists, tot_t = 0, 0.0
        for i in range(args.steps):
            n, t = step(100 + i)
            tot_lists += n
            tot_t += t
    value = tot_lists / tot_t
    sample = "%d processes x 1 image %dx%d x %d lists per step (K=%d), oracle port of sampling.py + NumPy ListMLE" % (
        cores, H, W, lists, K)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, json.dumps(shape)), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pldepth_b200 import _lib, synth
    from pldepth_b200.step import FusedPLStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # pinned host buffers of the e2e leg should live on the GPU's own NUMA node (first touch)
    from pldepth_b200.hostbind import bind_to_gpu_numa
    numa = ({"bound": False, "why": "PLD_NUMA_BIND=0"} if os.environ.get("PLD_NUMA_BIND", "1") == "0"
            else bind_to_gpu_numa(local_rank))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shape = workload_shape(args.workload)
    B, H, W, K, R = shape["B"], shape["H"], shape["W"], shape["K"], shape["R"]
    HW, L = H * W, B * R
    n_sets = max(1, args.sets)
    cfg_id = {"C1": 1, "C2": 2, "C3": 3, "C5s": 5, "C5": 5}[args.workload]

    # synthetic maps: a few distinct rank-transformed fields, rolled to make B distinct images
    base_maps = [synth.depth_map(H, W, 1000 * cfg_id + 17 * rank + i) for i in range(min(B, 4))]
    gt_h = np.stack([np.roll(base_maps[b % len(base_maps)], 31 * b, axis=1) for b in range(B)])
    mask_h = np.stack([synth.valid_mask(H, W, 3000 * cfg_id + b, args.hole) for b in range(B)])
    rs = np.random.RandomState(2000 * cfg_id + rank)
    pred_h = rs.standard_normal((B, H, W, 1)).astype(np.float32)

    sets = []
    for s in range(n_sets):
        sets.append(dict(gt=torch.from_numpy(np.roll(gt_h, s, axis=0)).to(dev),
                         mask=torch.from_numpy(mask_h).to(dev),
                         pred=torch.from_numpy(np.roll(pred_h, s, axis=0)).to(dev),
                         out=FusedPLStep.new_buffers(B, H, W, H, W, R, K, dev, emit_rankings=not args.no_emit)))
    from pldepth_b200._lib import Context
    step = FusedPLStep(K, R, seed=cfg_id, global_batch=B * world, image_base=rank * B, emit_rankings=not args.no_emit)
    # lanes: consecutive steps work on different buffer sets and do not depend on each other, so `lanes` of them
    # are kept in flight on separate streams.  Lane 0 is the plain step on the current stream.
    want_lanes = args.lanes if args.lanes is not None else (3 if L >= 100000 else 1)
    n_lanes = 1 if args.graph else max(1, min(want_lanes, n_sets))
    while n_sets % n_lanes:      # a buffer set must always be used by the same lane (stream order protects it)
        n_lanes -= 1
    lane_steps, lane_streams = [step], [torch.cuda.current_stream(dev)]
    for l in range(1, n_lanes):
        lane_steps.append(FusedPLStep(K, R, seed=cfg_id, global_batch=B * world, image_base=rank * B,
                                      emit_rankings=not args.no_emit, context=Context(local_rank),
                                      first_step=l << 24))
        lane_streams.append(torch.cuda.Stream(dev))

    pending = []

    def one_step(i, lanes=1):
        s = sets[i % n_sets]
        lane = i % lanes
        with torch.cuda.stream(lane_streams[lane]):
            out = lane_steps[lane].run(s["gt"], s["mask"], s["pred"], out=s["out"])
            if world > 1:
                # the path's only exchange: one f64 per step.  Nothing downstream of the step depends on it
                # (the gradient already carries the global 1/L), so it is issued asynchronously and overlaps
                # the next step; every reduction is waited for before the timed region closes.
                pending.append(dist.all_reduce(out["loss_sum"], async_op=True))
                if len(pending) > n_sets - 1:
                    pending.pop(0).wait()
        return out

    def fork_lanes(ev):
        for st in lane_streams[1:]:
            st.wait_event(ev)

    def join_lanes():
        cur = torch.cuda.current_stream(dev)
        for st in lane_streams[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    def drain():
        while pending:
            pending.pop(0).wait()

    for i in range(max(3, args.warmup)):
        one_step(i, n_lanes)
    drain()
    torch.cuda.synchronize()
    for ls in lane_steps:
        ls.check(dev)

    graphs = None
    if args.graph and world == 1:
        from pldepth_b200._lib import Context as _Ctx
        _Ctx.current(local_rank).device_offset(True, step.step_index)   # fresh draws on every replay
        graphs = []
        for s in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                one_step(s)
            graphs.append(g)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- dominant kernel alone: the library records CUDA events around the list kernel of every
    # step (pld_ctx_kernel_timing) on its launch stream, over a strictly sequential pass of the same steps
    # (one lane: nothing else runs beside the kernel), taken BEFORE the timed region so that the power
    # state left behind by the multi-lane burst cannot leak into it -----
    ctx = Context.current(local_rank)
    n_k = max(5, min(args.steps, 50))
    ctx.kernel_timing(n_k)
    for i in range(n_k):
        one_step(i)
    drain()
    torch.cuda.synchronize()
    kms = ctx.kernel_times(n_k)
    ctx.kernel_timing(0)
    kms = sorted(kms)
    k_ms = sum(kms) / len(kms)

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.rows.clear()
    e0.record()
    fork_lanes(e0)
    for i in range(args.steps):
        if graphs is not None:
            graphs[i % n_sets].replay()
        else:
            one_step(i, n_lanes)
    join_lanes()
    drain()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    if graphs is not None:
        launches = args.steps * 3
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    # clocks under load: nvidia-smi needs ~50-100 ms per query and the timed region lasts only a few ms, so the
    # SAME steps keep running for ~0.7 s right after it while the sampler polls.  The burst length is derived
    # from the max-reduced time, hence identical on every rank (each step carries a collective).
    n_burst = max(args.steps, min(20000, int(700.0 / max(ms_total / args.steps, 1e-3))))
    for i in range(n_burst):
        if graphs is not None:
            graphs[i % n_sets].replay()
        else:
            one_step(i, n_lanes)
    drain()
    torch.cuda.synchronize()
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    for ls in lane_steps:
        ls.check(dev)

    peak, peak_src = peaks()
    abytes = algorithmic_bytes(B, HW, L, K)
    achieved = abytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass

    # ---- end to end through the public API with HOST buffers ---------------------------------
    # HostPipelinedStep: every step copies its gt / mask / pred from pinned host memory, runs the
    # fused step and copies loss + dense gradient back; copies of neighbouring steps overlap kernels.
    from pldepth_b200.step import HostPipelinedStep
    gt_p = [torch.from_numpy(np.roll(gt_h, s, axis=0)).pin_memory() for s in range(2)]
    mask_p = torch.from_numpy(mask_h).pin_memory()
    pred_p = [torch.from_numpy(np.roll(pred_h, s, axis=0)).pin_memory() for s in range(2)]
    runner = HostPipelinedStep(K, R, B, H, W, seed=cfg_id, global_batch=B * world, image_base=rank * B,
                               emit_rankings=not args.no_emit)

    def e2e_step(i):
        t = runner.submit(gt_p[i % 2], mask_p, pred_p[i % 2])
        if world > 1:
            dist.all_reduce(runner.slots[t % 2]["out"]["loss_sum"])
        return t

    n_e2e = max(3, min(args.steps, 20))
    for i in range(3):
        e2e_step(i)
    barrier()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    last = 0
    for i in range(n_e2e):
        last = e2e_step(i)
    loss_host, grad_host = runner.result(last)          # waits for the final D2H
    b_.record()
    barrier()
    ms_e2e = a.elapsed_time(b_)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    h2d, d2h = runner.bytes_per_step()
    e2e = {"value": L * world * n_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / n_e2e, "steps": n_e2e,
           "api": "HostPipelinedStep.submit/result: pinned host gt+mask+pred -> device, fused step, loss + dense "
                  "gradient -> pinned host; 2 slots, H2D / compute / D2H streams overlap across steps",
           "last_loss": loss_host, "host_numa": numa}

    if rank == 0:
        line = {
            "metric": METRIC, "value": L * world * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s per GPU: B=%d images %dx%d, ranking_size K=%d, R=%d lists/image, %s "
                                   "mask, core sampler (factor 1.0), rankings %s" % (
                                       args.workload, B, H, W, K, R,
                                       "all-ones" if args.hole == 0 else "%.0f%%-hole" % (100 * args.hole),
                                       "not materialised" if args.no_emit else "emitted"),
                       "lists_per_step": L * world, "sharding": "per image, %d GPU(s)" % world,
                       "cache": "rotating %d input/output buffer sets of %.0f MB each (> 126 MB L2), no reuse "
                                "between consecutive steps" % (n_sets, abytes / 1e6),
                       "cuda_graph": bool(graphs),
                       "lanes": "%d independent batches in flight on separate streams (own lookup tables each); "
                                "roofline.kernel_ms is timed in a separate sequential pass" % n_lanes},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": ("lists_small_kernel<K,PHILOX_TAB,LOSS>" if K <= 16 else
                                    "lists_large_kernel<LPL,IPL,PHILOX_TAB,LOSS>") +
                                   " (fused sample+order+emit+gather+loss+bwd)",
                         "kernel_ms": k_ms, "algorithmic_bytes": abytes, "peak_source": peak_src},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_rate_single(shape, args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
