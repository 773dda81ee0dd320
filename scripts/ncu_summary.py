#!/usr/bin/env python
"""Per-kernel digest of an `ncu --set full` report: duration, DRAM bytes, DRAM/SM/tensor utilisation, launch shape.
Capture with a kernel filter so that torch's own initialisation kernels are not replayed under --set full:
  ncu --set full --clock-control none --import-source on -k regex:"tc_|bn_|dw_" -o rep python scripts/ncu_ops.py
Usage: python scripts/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        print("--- " + r[name_col][:110])
        for m in METRICS:
            if m in hdr:
                j = hdr.index(m)
                print(f"  {m:66s} {r[j]} {units[j]}")


if __name__ == "__main__":
    main()
