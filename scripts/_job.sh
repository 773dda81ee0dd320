#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_x2_gpu.py -q -x -k "dwconv or tc_gemm_x2" 2>&1 | tail -15
timeout 300 python scripts/bench_ops.py dwbwd 2>&1 | tee gpurun_out/dwbwd.log
timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | tee gpurun_out/x2gemm.log
