#!/bin/bash
mkdir -p gpurun_out
K='test_dwconv_fwd_stats_x2 and (case0 or case3 or case6 or case9) or test_dwconv_bwd_fused and (case0 or case3 or case10 or case13) or test_tc_gemm_x2 and (40000 or 1000)'
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 python -m pytest tests/test_kernels_gpu.py tests/test_x2_gpu.py -q -x -k "$K" -p no:cacheprovider > gpurun_out/sanitizer_new_${tool}.log 2>&1
  echo "== $tool rc=$?: $(grep -E ' passed| failed' gpurun_out/sanitizer_new_${tool}.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_new_${tool}.log | tail -1)"
done
