#!/bin/bash
mkdir -p gpurun_out
timeout 140 python bench.py --steps 5 --warmup 3 --modality rgb,sound,flow,rgbdiff --batch 48 --recompute --u8-input --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_cfg4_N48_1gpu.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_cfg4_N48_1gpu.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['peak_mem_gib'], d['gpu_launches'])" || tail -c 1500 gpurun_out/r2_bench_cfg4_N48_1gpu.log
