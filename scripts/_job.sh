#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_" -o gpurun_out/r2_prof_mid_after -f python scripts/ncu_mid.py > gpurun_out/ncu_mid_after.log 2>&1; tail -2 gpurun_out/ncu_mid_after.log
