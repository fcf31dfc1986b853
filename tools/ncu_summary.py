#!/usr/bin/env python
"""Summarise ncu outputs: launch list (csv from --metrics gpu__time_duration.sum) and key raw
metrics of an .ncu-rep (needs `ncu` on PATH; works without a GPU)."""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki][:90]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("%5s %10s %7s  kernel" % ("n", "avg_us", "share"))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%5d %10.1f %6.1f%%  %s" % (len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot, k))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:100])
        for k in KEYS:
            if k in h:
                print("   %-75s %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))


if __name__ == "__main__":
    for a in sys.argv[1:]:
        (raw if a.endswith(".ncu-rep") else launches)(a)
