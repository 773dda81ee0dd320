#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py -q -x -k "tc_gemm_x2 or tc_conv_x2 or golden" 2>&1 | tail -3 | cut -c1-250
for p in 0 1; do echo "== WIDE=$p"; ADAMML_B200_X2_WIDE=$p timeout 300 python scripts/bench_ops.py x2gemm 2>&1 | grep -E "K=256|K=512|K=144"; ADAMML_B200_X2_WIDE=$p timeout 300 python scripts/bench_ops.py x2conv 2>&1 | tail -4; done
