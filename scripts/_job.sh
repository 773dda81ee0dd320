#!/bin/bash
# scratch GPU job (run as: gpurun -- 'bash scripts/_job.sh')
mkdir -p gpurun_out
timeout 500 python bench.py --steps 5 --warmup 3 --modality rgb,flow,rgbdiff --batch 64 --recompute --u8-input --no-cpu-baseline > gpurun_out/r2_bench_cfg3_rgb_flow_N64_1gpu.log 2> gpurun_out/cfg3.err; tail -c 300 gpurun_out/cfg3.err; cut -c1-200 gpurun_out/r2_bench_cfg3_rgb_flow_N64_1gpu.log; grep -o '"peak_mem_gib": [0-9.]*' gpurun_out/r2_bench_cfg3_rgb_flow_N64_1gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 --modality rgb,sound,flow,rgbdiff --batch 48 --recompute --u8-input --no-cpu-baseline > gpurun_out/r2_bench_cfg4_N48_1gpu.log 2> gpurun_out/cfg4.err; tail -c 300 gpurun_out/cfg4.err; cut -c1-200 gpurun_out/r2_bench_cfg4_N48_1gpu.log; grep -o '"peak_mem_gib": [0-9.]*' gpurun_out/r2_bench_cfg4_N48_1gpu.log
