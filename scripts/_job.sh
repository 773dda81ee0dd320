#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py tests/test_kernels_gpu.py tests/test_blocks_gpu.py -q -x 2>&1 | tail -3 | cut -c1-250
timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | tail -13
timeout 300 python scripts/bench_ops.py shallow 2>&1 | tail -11
