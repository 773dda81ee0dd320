#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_x2_gpu.py tests/test_kernels_gpu.py tests/test_fused_eval_gpu.py tests/test_eval_skip_gpu.py -q -k "tc_ or fused or bn_act or first_conv or skip or selected or rng" 2>&1 | tail -3
timeout 300 python scripts/bench_ops.py x2gemm > gpurun_out/ops_x2gemm_v4.log 2>&1; cat gpurun_out/ops_x2gemm_v4.log
