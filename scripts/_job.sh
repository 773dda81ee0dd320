#!/bin/bash
mkdir -p gpurun_out
for m in 1 2 3; do echo "== MINB=$m"; ADAMML_B200_POOL_MINB=$m timeout 600 python -m pytest tests/test_x2_gpu.py -q -x -k "maxpool or stem_pool" 2>&1 | tail -1 | cut -c1-200; ADAMML_B200_POOL_MINB=$m timeout 300 python scripts/bench_pool.py 2>&1 | tail -2; done
