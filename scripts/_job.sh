#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py tests/test_kernels_gpu.py tests/test_blocks_gpu.py -q -x -k "bn or mask_bits or block" 2>&1 | tail -3 | cut -c1-250
for f in 1 0; do ADAMML_B200_MASK_BITS=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('mask_bits=$f', d['ms_per_step'], d['value'], d['config']['peak_mem_gib'], d['kernel_breakdown_ms'].get('bn_bwd_reduce'), d['kernel_breakdown_ms'].get('bn_apply_x2'))"; done
