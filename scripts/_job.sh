#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py tests/test_eval_skip_gpu.py tests/test_fused_eval_gpu.py tests/test_kernels_gpu.py -q -x -k "pool or skip or fused or eval" 2>&1 | tail -2 | cut -c1-200
timeout 300 python scripts/bench_pool.py 2>&1 | tail -2
