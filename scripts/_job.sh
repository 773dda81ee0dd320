#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_x2_gpu.py -q -x -k "bn" 2>&1 | tail -3
echo "--- one wave (new default)"; timeout 300 python scripts/bench_ops.py bnbwdred 2>&1 | tee gpurun_out/bnred_new.log
echo "--- 4 blocks per SM (old)"; ADAMML_B200_BN_REDUCE_BPSM=4 timeout 300 python scripts/bench_ops.py bnbwdred 2>&1 | tee gpurun_out/bnred_old.log
echo "--- 3 per SM"; ADAMML_B200_BN_REDUCE_BPSM=3 timeout 300 python scripts/bench_ops.py bnbwdred 2>&1 | tee gpurun_out/bnred_3.log
