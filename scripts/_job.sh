#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_2gpu.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.log').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['sync_bn'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d.get('latency_bound_ms'))"
timeout 600 python scripts/check_syncbn_p2p.py 2>&1 | tail -3
