#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2_bench_2gpu.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu.log').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['sync_bn'], d['gpu_launches'], d['roofline']['kernel'], round(d['roofline']['frac'],3))" || tail -c 2000 gpurun_out/r2_bench_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_bench_2gpu_nograph.log
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_nograph.log').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['sync_bn'], d['config']['cuda_graph'])" || tail -c 2000 gpurun_out/r2_bench_2gpu_nograph.log
