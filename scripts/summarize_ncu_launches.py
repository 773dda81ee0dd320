#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total time, share, launches.
Usage: python scripts/summarize_ncu_launches.py gpurun_out/launches.csv "<header comment>" > profiles/<name>.csv"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    m = re.search(r"(?:<unnamed>::|at::native::|\s)([A-Za-z_][A-Za-z0-9_]*)\s*(<[^(]*>)?\(", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    targs = re.sub(r"__nv_bfloat16", "bf16", targs)
    return base + (targs if len(targs) < 40 else "")


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e6 if r["Metric Unit"] in ("ns", "nsecond") else v / 1e3 if r["Metric Unit"] in ("us", "usecond") else v
            rows.append((short(r["Kernel Name"]), v))
    agg = defaultdict(lambda: [0.0, 0])
    for k, v in rows:
        agg[k][0] += v
        agg[k][1] += 1
    total = sum(v[0] for v in agg.values())
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    print(f"# total {total:.1f} ms over {len(rows)} launches; per-launch times are cold-cache and serialised: compare SHARES")
    print("kernel,total_ms,share_pct,launches")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"\"{k}\",{t:.2f},{100 * t / total:.1f},{n}")


if __name__ == "__main__":
    main()
