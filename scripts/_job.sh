#!/bin/bash
mkdir -p gpurun_out
for ph in main policy; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --phase $ph 2>&1 | tail -1 > gpurun_out/r2_bench_1gpu_phase_$ph.log; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu_phase_$ph.log').read()); print('$ph', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['peak_mem_gib'], d['config']['trainable'], d['roofline']['kernel'], round(d['roofline']['frac'],3))" || tail -5 gpurun_out/r2_bench_1gpu_phase_$ph.log; done
