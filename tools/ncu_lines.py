#!/usr/bin/env python
"""Attribute executed instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines.

  python tools/ncu_lines.py report.ncu-rep object.o 'mangled-kernel-substring' [top]

ncu's CSV export carries per-SASS-instruction counters but no source lines; `nvdisasm -g` of the same cubin carries
the line table.  Both list the kernel's instructions in the same order, so they are joined by position."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    data = rows[2:]
    ia, ism, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith("//---") and ".text." in l and kern in l][0]
    cur, seq = ("?", 0), []
    for l in lines[start + 1:]:
        if l.startswith("//---") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append(cur)
    if len(seq) != len(data):
        print("warning: %d disassembled instructions vs %d profiled" % (len(seq), len(data)))
    inst, smp = defaultdict(int), defaultdict(int)
    for loc, r in zip(seq, data):
        inst[loc] += int(r[ia])
        smp[loc] += int(r[ism])
    ti, ts = sum(inst.values()), sum(smp.values())
    print("total warp instructions %d, samples %d" % (ti, ts))
    src_cache = {}
    for loc, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        f, ln = loc
        text = ""
        for d in (os.path.dirname(os.path.abspath(obj)) + "/..", os.path.dirname(os.path.abspath(obj)), "."):
            p = os.path.join(d, f)
            if os.path.isfile(p):
                src_cache.setdefault(p, open(p).read().splitlines())
                if ln - 1 < len(src_cache[p]):
                    text = src_cache[p][ln - 1].strip()[:90]
                break
        print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * n / ti, 100.0 * smp[loc] / max(ts, 1), f, ln, text))


if __name__ == "__main__":
    main()
