#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_$i.log; tail -1 gpurun_out/pytest_$i.log | cut -c1-200; grep -E "^FAILED|^E " gpurun_out/pytest_$i.log | head -5 | cut -c1-300; done
cp gpurun_out/pytest_3.log gpurun_out/r2_pytest_gpu.log
