#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_tail_gpu.py -q -x 2>&1 | tail -2 | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --clip-gradient 20 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('clip', d['ms_per_step'], d['value'], d['gpu_launches'])"
