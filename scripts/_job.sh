#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_tail_gpu.py -q 2>&1 | tail -15
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x2_fusedtail.log 2> gpurun_out/ft.err; tail -c 300 gpurun_out/ft.err; cut -c1-200 gpurun_out/bench_x2_fusedtail.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --torch-tail > gpurun_out/bench_x2_torchtail.log 2> gpurun_out/tt.err; tail -c 300 gpurun_out/tt.err; cut -c1-200 gpurun_out/bench_x2_torchtail.log
