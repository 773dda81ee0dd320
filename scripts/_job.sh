#!/bin/bash
mkdir -p gpurun_out
bash scripts/run_sanitizers.sh 400 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --dump-calls gpurun_out/r2_calls_N72_x2.jsonl 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['peak_mem_gib'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['clocks']); print(d['kernel_breakdown_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-120
