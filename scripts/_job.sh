#!/bin/bash
# scratch GPU job (run as: gpurun -- 'bash scripts/_job.sh')
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_x2_gpu.py -q -k "dwconv or inverted" 2>&1 | tail -3
timeout 200 python scripts/bench_ops.py dw > gpurun_out/ops_dw_ffma2.log 2>&1; cat gpurun_out/ops_dw_ffma2.log
