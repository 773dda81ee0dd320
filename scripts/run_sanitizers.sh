#!/bin/bash
# compute-sanitizer over the tensor-core kernels (csrc/conv_tc.cu incl. the x2 path and the fused inference epilogue,
# csrc/conv_tc_wgrad.cu), the x2 row-streaming / depthwise / pooling kernels, the device-side gating kernels and one
# whole-model default-mode step, through the same pytest cases that check their numerics.
#   memcheck  = out-of-bounds / misaligned accesses      racecheck = shared-memory hazards between warp roles
#   initcheck = reads of uninitialised global memory (the gated inference pass reads only the live prefix)
#   bash scripts/run_sanitizers.sh [seconds per tool]   (GPU box; logs gpurun_out/sanitizer_*.log, summary on stdout)
# With 2 GPUs the peer-memory sync-BN exchange (csrc/p2p.cu) is covered by
#   compute-sanitizer --tool memcheck python -m torch.distributed.run --nproc-per-node 2 scripts/check_syncbn_p2p.py
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-300}
K1='test_tc_gemm_x2 and (1000 or 129 or 513) or test_tc_conv_x2 and (case0 or case3) or test_first_conv_x2 and 3-7-64 or test_tc_wgrad and case0 or bn_apply_and_stats_x2 or dwconv_fwd_x2 or test_dwconv_fwd_stats_x2 and (case0 or case3 or case6 or case9) or test_dwconv_bwd_fused and (case0 or case1 or case3 or case10 or case13) or test_tc_gemm_x2 and 40000 or pools_x2 or test_conv_bn_act_matches_unfused and case1 or test_bn_act_maxpool_x2 and (case0 or case1) or test_tc_conv_x2 and case4 or test_tc_gemm_x2 and 20002 or test_fused_clip_grad_norm'
FILES="tests/test_kernels_gpu.py tests/test_x2_gpu.py tests/test_fused_eval_gpu.py tests/test_train_tail_gpu.py"
for tool in memcheck racecheck; do
  log=gpurun_out/sanitizer_${tool}.log
  timeout $T compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 \
    python -m pytest $FILES -q -x -k "$K1" -p no:cacheprovider > $log 2>&1
  echo "== $tool rc=$?: $(grep -E ' passed| failed' $log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done
# device-side gating + one whole-model default-mode (x2) training step under memcheck
timeout $T compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 \
  python -m pytest tests/test_eval_skip_gpu.py tests/test_x2_gpu.py -q -x \
  -k "test_rng_policy_extremes and device and 0.5 or default_mode_matches_reference_golden and nocausal" -p no:cacheprovider \
  > gpurun_out/sanitizer_memcheck_model.log 2>&1
echo "== memcheck gating + whole model rc=$?: $(grep -E ' passed| failed' gpurun_out/sanitizer_memcheck_model.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_model.log | tail -1)"
