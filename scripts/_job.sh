#!/bin/bash
mkdir -p gpurun_out
for i in $(seq 1 14); do timeout 300 python -m pytest tests/test_blocks_gpu.py -q -x -k "test_resnet_block" --tb=short 2>&1 | grep -E "^E |passed|failed" | head -8 | cut -c1-250; done
