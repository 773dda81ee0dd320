#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_x2_gpu.py -q -x -k "pack_cache" --tb=short 2>&1 | tail -12 | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['host_issue_ms_per_step_eager'], d['clocks']['sm_mhz'], d['config']['peak_mem_gib'])"
