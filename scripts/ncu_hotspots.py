#!/usr/bin/env python
"""Top warp-stall sampling hot spots (SASS, with a few preceding instructions for context) of an ncu report.
Usage: python scripts/ncu_hotspots.py report.ncu-rep [min_share=0.04]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.04
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kernels = out.split('"Address","Source"')
    for kidx, chunk in enumerate(kernels[1:]):
        rows = list(csv.reader(io.StringIO('"Address","Source"' + chunk)))
        H = rows[0]
        cs, j, ie = H.index("Source"), H.index("# Samples"), H.index("Instructions Executed")
        body = [r for r in rows[1:] if len(r) > max(cs, j, ie)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        tot = sum(num(r[j]) for r in body) or 1.0
        print(f"==== kernel #{kidx}: {int(tot)} samples")
        for h, r in enumerate(body):
            if num(r[j]) / tot >= min_share:
                print("  ----")
                for i in range(max(0, h - 8), min(len(body), h + 2)):
                    q = body[i]
                    print(f"  {i:5d} {100 * num(q[j]) / tot:5.1f}% ex={q[ie]:>9} {q[cs][:96]}")


if __name__ == "__main__":
    main()
