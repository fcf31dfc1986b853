#!/usr/bin/env python
"""Print the key numbers of bench.py JSON lines read from stdin or files: one short row per line."""
import json
import sys


def rows(stream):
    for ln in stream:
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                yield json.loads(ln)
            except ValueError:
                pass


def show(tag, d):
    r = d.get("roofline", {})
    e = d.get("e2e", {})
    print("%-28s %.4g %s  ms/step %.4f  kernel_ms %s  frac %s  step_frac %s  e2e ms %s" % (
        tag, d.get("value", 0), d.get("unit", ""), d.get("ms_per_step", 0), r.get("kernel_ms"), r.get("frac"),
        r.get("step_frac"), e.get("ms_per_step")))
    for s in d.get("secondary", []) or []:
        print("   secondary %-40s ms/step %.4f kernel_ms %s frac %s" % (s.get("workload"), s.get("ms_per_step", 0),
                                                                      s.get("kernel_ms"), s.get("frac")))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for f in sys.argv[1:]:
            for d in rows(open(f)):
                show(f, d)
    else:
        for d in rows(sys.stdin):
            show("-", d)
