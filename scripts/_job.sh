#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --dump-calls gpurun_out/r2_calls_N72_x2.jsonl 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu_nograph.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu_bf16.log
timeout 600 python bench.py --steps 5 --warmup 3 --recompute --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu_recompute.log
for f in r2_bench_1gpu r2_bench_1gpu_nograph r2_bench_1gpu_bf16 r2_bench_1gpu_recompute; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.log').read()); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['peak_mem_gib'], d['roofline']['kernel'], round(d['roofline']['frac'],3))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
ls -la gpurun_out/launches.csv
