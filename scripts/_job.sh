#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py -q -x -k "tc_gemm_x2 or tc_conv_x2 or golden" 2>&1 | tail -3 | cut -c1-250
for p in 0 1; do echo "== SPLIT=$p"; ADAMML_B200_TC_SPLIT=$p timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | tail -13; done
