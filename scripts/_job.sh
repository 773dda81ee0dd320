#!/bin/bash
# scratch GPU job (run as: gpurun -- 'bash scripts/_job.sh')
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_x2_gpu.py -q -s -k "recompute" 2>&1 | tail -5
timeout 400 python bench.py --steps 6 --warmup 3 --recompute --no-cpu-baseline > gpurun_out/bench_x2_recompute.log 2> gpurun_out/bench_x2_recompute.err; tail -c 300 gpurun_out/bench_x2_recompute.err; cut -c1-330 gpurun_out/bench_x2_recompute.log; grep -o '"peak_mem_gib": [0-9.]*' gpurun_out/bench_x2_recompute.log
