#!/bin/bash
mkdir -p gpurun_out
bash scripts/run_sanitizers.sh 240 2>&1 | tee gpurun_out/r2_sanitizer_summary.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2_bench_1gpu_nograph.log 2> gpurun_out/nograph.err; tail -c 200 gpurun_out/nograph.err; cut -c1-220 gpurun_out/r2_bench_1gpu_nograph.log
