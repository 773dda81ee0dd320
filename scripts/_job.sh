#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
timeout 300 python scripts/bench_ops.py dwfwd 2>&1 | tee gpurun_out/dwfwd.log | tail -4
timeout 600 python bench.py 2>&1 | tail -2 > gpurun_out/bench_a.log; tail -c 3000 gpurun_out/bench_a.log
