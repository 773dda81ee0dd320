#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_x2_gpu.py -q -k "bn_apply_and_stats" 2>&1 | tail -2
timeout 200 python scripts/bench_ops.py x2bn 2>&1
