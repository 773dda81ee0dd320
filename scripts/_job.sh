#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_eval_skip.py > gpurun_out/r2_eval_skip_N72_S10_device.log 2>&1; tail -8 gpurun_out/r2_eval_skip_N72_S10_device.log | cut -c1-600
