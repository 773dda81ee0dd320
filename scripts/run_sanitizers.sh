#!/bin/bash
# compute-sanitizer over the tensor-core kernels (csrc/conv_tc.cu incl. the x2 path, csrc/conv_tc_wgrad.cu), the
# row-streaming BN / depthwise / pool kernels and one whole-model x2 step, through the same pytest cases that check
# their numerics.  memcheck = out-of-bounds / misaligned accesses (incl. TMA-written shared memory), racecheck =
# shared-memory hazards between the producer / MMA / epilogue warps, initcheck = reads of uninitialised global memory.
#   bash scripts/run_sanitizers.sh   (on a GPU box; writes gpurun_out/sanitizer_*.log, summary to stdout)
# With 2 GPUs the peer-memory sync-BN exchange (csrc/p2p.cu) is covered by
#   compute-sanitizer --tool memcheck torchrun ... scripts/check_syncbn_p2p.py   (see profiles/README.md)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='tc_gemm or tc_conv or tc_wgrad or tc_stem or first_conv or bn_apply_and_stats_x2 or dwconv_fwd_x2 or pools_x2'
FILES="tests/test_kernels_gpu.py tests/test_x2_gpu.py"
for tool in memcheck racecheck initcheck; do
  log=gpurun_out/sanitizer_${tool}.log
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest $FILES -q -x -k "$SEL" -p no:cacheprovider > $log 2>&1
  rc=$?
  echo "== $tool rc=$rc: $(grep -E 'passed|failed' $log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done
# one whole-model default-mode (x2) training step under memcheck
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_x2_gpu.py -q -x -k "default_mode_matches_reference_golden and nocausal" -p no:cacheprovider \
  > gpurun_out/sanitizer_memcheck_model.log 2>&1
echo "== memcheck whole model rc=$?: $(grep -E 'passed|failed' gpurun_out/sanitizer_memcheck_model.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_model.log | tail -1)"
