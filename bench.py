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
    ap.add_argument("--strategy", default="purely", choices=["purely", "masked", "thresholded", "information"],
                    help="sampling strategy (default: the core sampler, exactly R lists per image)")
    ap.add_argument("--secondary", dest="secondary", action="store_true", default=None,
                    help="also measure the secondary workloads (default: on for the default C2 headline at N=1)")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    a = ap.parse_args()
    if a.secondary is None:
        a.secondary = (a.workload == "C2" and a.hole == 0.0 and a.strategy == "purely" and not a.no_emit and
                       a.gpus == 1 and a.impl == "b200")
    return a


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
    """Samples SM clocks / throttle reasons of the given GPUs during the timed region: in-process NVML every 20 ms
    (no process spawn: eight ranks forking nvidia-smi every 200 ms is a measurable host load at N = 8), falling back to
    the nvidia-smi command line where pynvml is missing.  Only rank 0 samples, for all GPUs of the job."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # NVML clocks-event-reason bits
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices = list(indices) if isinstance(indices, (list, tuple, range)) else [indices]
        self.rows = []            # [sm_mhz, max_mhz, power_w, hw, hw_thermal, sw_thermal, sw_power_cap] as strings
        self.stop_flag = threading.Event()
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        for h in self._handles:
            sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
            try:
                pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = int(get(h))
            self.rows.append([str(sm), str(mx), "%.1f" % pw] +
                             ["Active" if bits & b else "Not Active" for _, b in self.BITS])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self._nvml is not None else 0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source, "gpus": len(self.indices)}


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
            "config": {"workload": workload_string(args.workload, shape, args.hole, not args.no_emit, args.strategy),
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
SM_COUNT = 148
# micro-benchmark floor of the unit that binds the list kernels (profiles/r01_microbench.jsonl): per SM and clock one
# divergent 8-byte table gather, 0.66 scattered float reductions, one store sector
FLOOR_GATHER_CLK, FLOOR_RED_CLK, FLOOR_STORE_SECTOR_CLK = 1.0, 1.0 / 0.66, 1.0


def workload_string(name, shape, hole, emit, strategy):
    """One description of the measured configuration, identical in both arms."""
    return "%s per GPU: B=%d images %dx%d, ranking_size K=%d, R=%d lists/image, %s mask, %s, rankings %s" % (
        name, shape["B"], shape["H"], shape["W"], shape["K"], shape["R"],
        "all-ones" if hole == 0 else "%.0f%%-hole" % (100 * hole),
        "core sampler (factor 1.0)" if strategy == "purely" else "%s strategy (R best of its candidates)" % strategy,
        "emitted" if emit else "not materialised")


def synthetic_batch(name, shape, hole, rank):
    import numpy as np
    from pldepth_b200 import synth
    B, H, W = shape["B"], shape["H"], shape["W"]
    cfg_id = {"C1": 1, "C2": 2, "C3": 3, "C5s": 5, "C5": 5}[name]
    # a few distinct rank-transformed fields, rolled to make B distinct images
    base_maps = [synth.depth_map(H, W, 1000 * cfg_id + 17 * rank + i) for i in range(min(B, 4))]
    gt_h = np.stack([np.roll(base_maps[b % len(base_maps)], 31 * b, axis=1) for b in range(B)])
    mask_h = np.stack([synth.valid_mask(H, W, 3000 * cfg_id + b, hole) for b in range(B)])
    pred_h = np.random.RandomState(2000 * cfg_id + rank).standard_normal((B, H, W, 1)).astype(np.float32)
    return cfg_id, gt_h, mask_h, pred_h


class Measurement(object):
    """Device-resident measurement of one workload: rotating buffer sets, optional lanes, kernel-alone timing."""

    def __init__(self, name, shape, hole, emit, strategy, n_sets, want_lanes, world, rank, local_rank, graph=False):
        import numpy as np
        import torch
        from pldepth_b200._lib import Context
        from pldepth_b200.dist import LossWindow
        from pldepth_b200.step import FusedPLStep
        self.torch, self.np = torch, np
        self.name, self.shape, self.hole, self.emit, self.strategy = name, shape, hole, emit, strategy
        self.world, self.rank, self.local_rank = world, rank, local_rank
        dev = self.dev = torch.device("cuda", local_rank)
        B, H, W, K, R = shape["B"], shape["H"], shape["W"], shape["K"], shape["R"]
        self.L = B * R
        self.cfg_id, self.gt_h, self.mask_h, self.pred_h = synthetic_batch(name, shape, hole, rank)
        self.n_sets = n_sets = max(1, n_sets)
        self.sets = []
        for s in range(n_sets):
            self.sets.append(dict(gt=torch.from_numpy(np.roll(self.gt_h, s, axis=0)).to(dev),
                                  mask=torch.from_numpy(self.mask_h).to(dev),
                                  pred=torch.from_numpy(np.roll(self.pred_h, s, axis=0)).to(dev),
                                  out=FusedPLStep.new_buffers(B, H, W, H, W, R, K, dev, emit_rankings=emit)))
        n_lanes = 1 if graph else max(1, min(want_lanes, n_sets))
        while n_sets % n_lanes:      # a buffer set must always be used by the same lane (stream order protects it)
            n_lanes -= 1
        self.n_lanes = n_lanes
        mk = lambda l: FusedPLStep(K, R, seed=self.cfg_id, global_batch=B * world, image_base=rank * B,
                                   emit_rankings=emit, strategy=strategy,
                                   context=None if l == 0 else Context(local_rank), first_step=l << 24)
        self.lane_steps = [mk(l) for l in range(n_lanes)]
        self.lane_streams = [torch.cuda.current_stream(dev)] + [torch.cuda.Stream(dev) for _ in range(1, n_lanes)]
        # The path's only exchange -- the SUM of one float64 per step -- is taken off the step: every step writes its
        # local loss sum into a slot of its lane's window (the fused step takes the pointer), and a window is
        # all-reduced once per WINDOW steps.  Two windows per lane alternate, so a step never waits for a reduction.
        self.window_steps = 16
        self.windows = [[LossWindow(self.window_steps, dev), LossWindow(self.window_steps, dev)] for _ in range(n_lanes)]
        self.win_work = [[None, None] for _ in range(n_lanes)]
        self.lane_count = [0] * n_lanes
        self.reductions = 0
        self.last_losses = None

    def _close_window(self, lane, which):
        import torch.distributed as dist
        w = self.windows[lane][which]
        if w.filled == 0:
            return
        self.win_work[lane][which] = w.reduce(async_op=True) if self.world > 1 else w.reduce()
        self.reductions += 1 if self.world > 1 else 0
        self.last_losses = (lane, which)

    def one_step(self, i, lanes=None):
        torch = self.torch
        lanes = self.n_lanes if lanes is None else lanes
        s = self.sets[i % self.n_sets]
        lane = i % lanes
        with torch.cuda.stream(self.lane_streams[lane]):
            c = self.lane_count[lane]
            which = (c // self.window_steps) & 1
            w = self.windows[lane][which]
            if c % self.window_steps == 0 and self.win_work[lane][which] is not None:
                self.win_work[lane][which].wait()        # stream-level wait for this window's previous reduction
                self.win_work[lane][which] = None
            s["out"]["loss_sum"] = w.slot(c)
            out = self.lane_steps[lane].run(s["gt"], s["mask"], s["pred"], out=s["out"])
            self.lane_count[lane] = c + 1
            if w.mark():
                self._close_window(lane, which)
        return out

    def drain(self):
        """Reduce every partially filled window and wait (stream-level) for all outstanding reductions."""
        torch = self.torch
        for lane in range(self.n_lanes):
            with torch.cuda.stream(self.lane_streams[lane]):
                for which in (0, 1):
                    self._close_window(lane, which)
                    if self.win_work[lane][which] is not None:
                        self.win_work[lane][which].wait()
                        self.win_work[lane][which] = None
                # restart window accounting on a boundary
                self.lane_count[lane] = ((self.lane_count[lane] + self.window_steps - 1) // self.window_steps) * self.window_steps

    def fork_lanes(self, ev):
        for st in self.lane_streams[1:]:
            st.wait_event(ev)

    def join_lanes(self):
        cur = self.torch.cuda.current_stream(self.dev)
        for st in self.lane_streams[1:]:
            ev = self.torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)

    def barrier(self):
        import torch.distributed as dist
        self.drain()
        self.join_lanes()
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize()

    def check(self):
        for ls in self.lane_steps:
            ls.check(self.dev)

    def kernel_alone(self, n_k):
        """Mean duration of the dominant (list) kernel over a strictly sequential pass: the library records CUDA
        events around that launch on its stream (pld_ctx_kernel_timing).  Also returns the sequential step time."""
        from pldepth_b200._lib import Context
        torch = self.torch
        ctx = Context.current(self.local_rank)
        ctx.kernel_timing(n_k)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(n_k):
            self.one_step(i, lanes=1)
        b_.record()
        self.drain()
        torch.cuda.synchronize()
        kms = ctx.kernel_times(n_k)
        ctx.kernel_timing(0)
        # the sequential step time is taken in a second pass WITHOUT those event pairs: an event record between two
        # kernels turns their programmatic dependent launch back into a plain serialised one
        for i in range(2):
            self.one_step(i, lanes=1)
        torch.cuda.synchronize()
        a.record()
        for i in range(n_k):
            self.one_step(i, lanes=1)
        b_.record()
        self.drain()
        torch.cuda.synchronize()
        return sum(kms) / max(len(kms), 1), a.elapsed_time(b_) / n_k

    def timed(self, steps, graphs=None):
        import torch.distributed as dist
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        self.fork_lanes(e0)
        h0 = time.perf_counter()
        for i in range(steps):
            if graphs is not None:
                graphs[i % self.n_sets].replay()
            else:
                self.one_step(i)
        self.host_enqueue_ms = 1e3 * (time.perf_counter() - h0) / max(steps, 1)   # host time to enqueue one step
        self.drain()
        self.join_lanes()
        e1.record()
        self.barrier()
        ms_total = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total

    def roofline(self, k_ms, ms_per_step, seq_ms_per_step, peak, peak_src, sm_mhz):
        B, H, W, K, R = (self.shape[k] for k in "BHWKR")
        abytes = algorithmic_bytes(B, H * W, self.L, K) if self.emit else B * H * W * 16 + 4
        achieved = abytes / (k_ms * 1e-3) / 1e9
        clk = (sm_mhz or 1965.0) * 1e6
        cyc = k_ms * 1e-3 * clk * SM_COUNT / self.L
        floor = K * (FLOOR_GATHER_CLK + FLOOR_RED_CLK) + (K * 8.0 / 32.0) * FLOOR_STORE_SECTOR_CLK * (1 if self.emit else 0)
        if self.hole > 0 and self.emit and self.strategy == "purely":
            floor += K * FLOOR_GATHER_CLK       # holed + emitted: the prediction needs its own gather
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "step_frac": abytes / (ms_per_step * 1e-3) / 1e9 / peak,
                "step_frac_sequential": abytes / (seq_ms_per_step * 1e-3) / 1e9 / peak,
                "kernel": ("lists_small_kernel<K,PHILOX_TAB,LOSS>" if K <= 16 else
                           "lists_tab_kernel<LPL,IPL,THREADS,LOSS>") + " (fused sample+order+emit+gather+loss+bwd)",
                "kernel_ms": k_ms, "sequential_ms_per_step": seq_ms_per_step, "algorithmic_bytes": abytes,
                "peak_source": peak_src,
                "binding_unit": {"unit": "L1TEX sector operations (divergent gathers, reductions, store sectors)",
                                 "sm_cycles_per_list": cyc, "microbench_floor_cycles_per_list": floor,
                                 "frac_of_floor": floor / cyc,
                                 "source": "profiles/r01_microbench.jsonl: 1.0 gather, 0.66 reduction, 1 store sector "
                                           "per SM and clock; DESIGN.md section 4"}}

    def free(self):
        self.sets = None
        self.lane_steps = None
        self.torch.cuda.empty_cache()


def graph_replay_ms(m, shape, strategy, local_rank, iters=8):
    """ms per step of the workload of `m` (rankings not materialised) replayed from CUDA graphs, one per buffer set,
    captured on a private context so that the device-resident Philox offset stays out of everybody else's way."""
    import torch
    from pldepth_b200._lib import Context
    from pldepth_b200.step import FusedPLStep
    step = FusedPLStep(shape["K"], shape["R"], seed=m.cfg_id, global_batch=shape["B"] * m.world, image_base=m.rank * shape["B"],
                       emit_rankings=False, strategy=strategy, context=Context(local_rank))
    graphs = [step.capture(s["gt"], s["mask"], s["pred"])[0] for s in m.sets]
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    step.check(m.dev)
    return e0.elapsed_time(e1) / iters


def secondary_rows(world, rank, local_rank, peak, peak_src, sm_mhz):
    """The workloads the reference really runs, measured in the same process after the headline (sequential steps,
    few iterations): holed masks, long lists, the config-5 share, the default (InformationScore) strategy."""
    rows = []
    plan = [("C2", 0.1, True, "purely"), ("C2", 0.1, False, "purely"), ("C2", 0.0, False, "purely"),
            ("C3", 0.0, True, "purely"), ("C3", 0.0, False, "purely"), ("C5", 0.0, True, "purely"),
            ("C2", 0.0, True, "information"), ("C2", 0.0, False, "information"),
            ("C2", 0.0, True, "thresholded"), ("C2", 0.0, False, "thresholded"),
            # long lists with the score-based strategies (hyperopt/hyperparams.py:44 sweeps ranking_size to 500)
            ("C3", 0.0, False, "thresholded"), ("C3", 0.0, False, "information")]
    for name, hole, emit, strategy in plan:
        shape = workload_shape(name)
        try:
            m = Measurement(name, shape, hole, emit, strategy, 2, 1, world, rank, local_rank)
            for i in range(3):
                m.one_step(i)
            m.barrier()
            m.check()
            k_ms, seq_ms = m.kernel_alone(8)
            ms = m.timed(8) / 8
            r = m.roofline(k_ms, ms, seq_ms, peak, peak_src, sm_mhz)
            rows.append({"workload": workload_string(name, shape, hole, emit, strategy), "lists_per_step": m.L * world,
                         "value": m.L * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kernel_ms": k_ms,
                         "frac": r["frac"], "step_frac": r["step_frac"],
                         "sm_cycles_per_list": r["binding_unit"]["sm_cycles_per_list"],
                         "floor_cycles_per_list": r["binding_unit"]["microbench_floor_cycles_per_list"],
                         "note": "frac = algorithmic bytes / LAST list kernel of the step; scored strategies run two list "
                                 "passes plus selection, see step_frac" if strategy != "purely" else
                                 "frac = algorithmic bytes / list-kernel time"})
            if emit and hole == 0.0 and strategy == "purely":
                # DRAM bytes of one launch of the list kernel from an ncu --set full capture (profiles/traffic.json)
                try:
                    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                        rows[-1]["traffic"] = json.load(f).get(name)
                except Exception:
                    rows[-1]["traffic"] = None
            if strategy != "purely" and not emit:
                # the same step replayed from a CUDA graph (FusedPLStep.capture: Philox offset in device memory, fresh
                # lists on every replay): a scored step is a dozen launches, most of them a few microseconds long
                try:
                    rows[-1]["ms_per_step_cuda_graph"] = graph_replay_ms(m, shape, strategy, local_rank)
                except Exception as exc:
                    rows[-1]["cuda_graph_error"] = repr(exc)[:200]
            m.free()
            del m
        except Exception as exc:   # a secondary row must never cost the headline line
            rows.append({"workload": workload_string(name, shape, hole, emit, strategy), "error": repr(exc)[:300]})
    return rows


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pldepth_b200 import _lib
    from pldepth_b200.dist import LossWindow

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
    L = B * R
    emit = not args.no_emit
    want_lanes = args.lanes if args.lanes is not None else (3 if L >= 100000 else 1)
    m = Measurement(args.workload, shape, args.hole, emit, args.strategy, args.sets, want_lanes, world, rank,
                    local_rank, graph=args.graph)
    n_sets, n_lanes = m.n_sets, m.n_lanes

    for i in range(max(3, args.warmup)):
        m.one_step(i)
    m.barrier()
    m.check()

    graphs = None
    if args.graph and world == 1:
        from pldepth_b200._lib import Context as _Ctx
        _Ctx.current(local_rank).device_offset(True, m.lane_steps[0].step_index)   # fresh draws on every replay
        graphs = []
        for s in range(n_sets):
            m.sets[s]["out"]["loss_sum"] = m.windows[0][0].slot(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m.lane_steps[0].run(m.sets[s]["gt"], m.sets[s]["mask"], m.sets[s]["pred"], out=m.sets[s]["out"])
            graphs.append(g)
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()

    # ---- dominant kernel alone, over a strictly sequential pass taken BEFORE the timed region so that the power
    # state left behind by the multi-lane burst cannot leak into it -----
    n_k = max(5, min(args.steps, 50))
    k_ms, seq_ms = m.kernel_alone(n_k)

    # rank 0 samples the clocks of every GPU of the job (one process per GPU, GPU index == local rank)
    sampler = ClockSampler(list(range(world)) if rank == 0 else [])
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    red0 = m.reductions
    sampler.rows.clear()
    ms_total = m.timed(args.steps, graphs)
    host_enqueue_ms = m.host_enqueue_ms
    if world > 1:     # slowest host, like the device time
        t = torch.tensor([host_enqueue_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        host_enqueue_ms = float(t.item())
    launches = _lib.launch_count() - launches0
    reductions = m.reductions - red0
    if graphs is not None:
        launches = args.steps * 3
    # clocks under load: nvidia-smi needs ~50-100 ms per query and the timed region lasts only a few ms, so the
    # SAME steps keep running for ~0.7 s right after it while the sampler polls.  The burst length is derived
    # from the max-reduced time, hence identical on every rank (the windows reduce collectively).
    n_burst = max(args.steps, min(20000, int(700.0 / max(ms_total / args.steps, 1e-3))))
    for i in range(n_burst):
        if graphs is not None:
            graphs[i % n_sets].replay()
        else:
            m.one_step(i)
    m.barrier()
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=3)
    m.check()
    clocks = sampler.summary()

    peak, peak_src = peaks()
    roof = m.roofline(k_ms, ms_total / args.steps, seq_ms, peak, peak_src, clocks.get("sm_mhz"))
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload if emit and args.hole == 0 and args.strategy == "purely" else "-")
    except Exception:
        pass
    roof["traffic"] = traffic

    # ---- end to end through the public API with HOST buffers ---------------------------------
    # HostPipelinedStep: every step copies its gt / mask / pred from pinned host memory, runs the
    # fused step and copies loss + dense gradient back; copies of neighbouring steps overlap kernels.
    # At N > 1 the per-step loss sums go through a LossWindow: one all-reduce + one D2H of the reduced
    # window close the timed region, so the loss that reaches the host is the GLOBAL one.
    from pldepth_b200.step import HostPipelinedStep
    gt_h, mask_h, pred_h = m.gt_h, m.mask_h, m.pred_h
    cfg_id = m.cfg_id
    n_e2e = max(3, min(args.steps, 20))

    def e2e_leg(mask_dtype):
        gt_p = [torch.from_numpy(np.roll(gt_h, s, axis=0)).pin_memory() for s in range(2)]
        mask_src = mask_h.astype(np.uint8) if mask_dtype == torch.uint8 else mask_h
        mask_p = torch.from_numpy(mask_src).pin_memory()
        pred_p = [torch.from_numpy(np.roll(pred_h, s, axis=0)).pin_memory() for s in range(2)]
        runner = HostPipelinedStep(K, R, B, H, W, seed=cfg_id, global_batch=B * world, image_base=rank * B,
                                   emit_rankings=emit, mask_dtype=mask_dtype)
        window = LossWindow(n_e2e, dev)
        h_window = torch.empty(n_e2e, dtype=torch.float64).pin_memory()

        def e2e_step(i, record):
            sl = runner.slots[runner.count % len(runner.slots)]
            if record:
                sl["out"]["loss_sum"] = window.slot(i)
            return runner.submit(gt_p[i % 2], mask_p, pred_p[i % 2])

        for i in range(3):
            e2e_step(i, False)
        m.barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        last = 0
        for i in range(n_e2e):
            last = e2e_step(i, True)
            window.mark()
        loss_host, grad_host = runner.result(last)          # waits for the final D2H (local loss + dense gradient)
        window.reduce()                                     # N > 1: the one collective of the leg
        h_window.copy_(window.values(), non_blocking=True)
        b_.record()
        torch.cuda.synchronize()
        if world > 1:
            loss_host = float(h_window[-1]) / float(B * world * R)
        m.barrier()
        ms = a.elapsed_time(b_)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        h2d, d2h = runner.bytes_per_step()
        return {"value": L * world * n_e2e / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h) + 8, "ms_per_step": ms / n_e2e, "steps": n_e2e,
                "last_loss": loss_host, "loss_is_global": True}

    e2e = e2e_leg(torch.float32)
    e2e["api"] = ("HostPipelinedStep.submit/result: pinned host gt+mask+pred (float32) -> device, fused step, loss + "
                  "dense gradient -> pinned host; 2 slots, H2D / compute / D2H streams overlap across steps; at N>1 "
                  "the per-step loss sums are all-reduced once (LossWindow) and read back inside the timed region")
    e2e["host_numa"] = numa
    e2e_u8 = None
    if args.strategy == "purely":
        e2e_u8 = e2e_leg(torch.uint8)
        e2e_u8["api"] = "same, mask handed over as uint8 (nonzero = valid; pld_fused_step_m8): a quarter of the mask bytes"

    secondary = None
    if world == 1 and args.secondary:
        m.free()
        secondary = secondary_rows(world, rank, local_rank, peak, peak_src, clocks.get("sm_mhz"))

    if rank == 0:
        line = {
            "metric": METRIC, "value": L * world * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, shape, args.hole, emit, args.strategy),
                       "lists_per_step": L * world, "sharding": "per image, %d GPU(s)" % world,
                       "cache": "rotating %d input/output buffer sets of %.0f MB each (> 126 MB L2), no reuse "
                                "between consecutive steps" % (n_sets, roof["algorithmic_bytes"] / 1e6),
                       "cuda_graph": bool(graphs),
                       "host_enqueue_ms_per_step": host_enqueue_ms,
                       "lanes": "%d independent batches in flight on separate streams (own lookup tables each); "
                                "roofline.kernel_ms is timed in a separate sequential pass" % n_lanes,
                       "collective": "loss sums all-reduced once per %d steps per lane (LossWindow), %d all-reduce(s) "
                                     "inside the timed region" % (m.window_steps, reductions)},
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e_u8 is not None:
            line["e2e_u8_mask"] = e2e_u8
        if secondary is not None:
            line["secondary"] = secondary
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
