#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_x2_gpu.py -q -x -k "pack_cache" --tb=short 2>&1 | tail -12 | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; tail -1 gpurun_out/r2_pytest_gpu.log | cut -c1-200; grep -E "^FAILED|^E " gpurun_out/r2_pytest_gpu.log | head -5 | cut -c1-300
