#!/usr/bin/env python
"""SASS opcode census of libadamml_b200.so: proof that the shipped kernels are Blackwell tensor-core / TMA code
(B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier).

    python scripts/sass_census.py > profiles/r2_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adamml_b200", "lib", "libadamml_b200.so")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "UTCATOMSWS", "HMMA", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    arch = set()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    per[cur][o] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print(f"library: {os.path.relpath(LIB, ROOT)}   arch: {sorted(arch)}   kernels: {len(per)}")
    print("total: " + "  ".join(f"{o}={total[o]}" for o in OPS))
    print()
    print(f"{'kernel (demangled prefix)':88s} " + " ".join(f"{o:>8s}" for o in OPS[:7]))
    for name, c in per.items():
        if not any(c[o] for o in OPS[:7]):
            continue
        try:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        except Exception:
            dem = name
        dem = dem.replace("(anonymous namespace)::", "")
        dem = re.sub(r"^void ", "", dem)
        dem = re.sub(r">\(.*", ">", dem) if "<" in dem else re.sub(r"\(.*", "", dem)
        print(f"{dem[:88]:88s} " + " ".join(f"{c[o]:8d}" for o in OPS[:7]))


if __name__ == "__main__":
    sys.exit(main())
