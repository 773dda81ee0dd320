#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; tail -1 gpurun_out/r2_pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --dump-calls gpurun_out/r2_calls_N72_x2.jsonl 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu_nograph.log
for f in r2_bench_1gpu r2_bench_1gpu_nograph; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.log').read()); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['peak_mem_gib'], d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-100
