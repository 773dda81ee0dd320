#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_x2_gpu.py -q -x -k "tc_gemm_x2" 2>&1 | tail -3
echo "--- fence (default)"; timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | tee gpurun_out/x2gemm_b.log
echo "--- no fence"; ADAMML_B200_TC_NOFENCE=1 timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | tee gpurun_out/x2gemm_nofence.log
ADAMML_B200_TC_NOFENCE=1 timeout 600 python -m pytest tests/test_x2_gpu.py -q -x -k "tc_gemm_x2 or tc_conv_x2" 2>&1 | tail -3
