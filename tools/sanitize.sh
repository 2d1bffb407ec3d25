#!/bin/bash
# compute-sanitizer passes over the small parity tests of every kernel family (run on a GPU box through gpurun):
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      default: memcheck racecheck
# Writes gpurun_out/sanitize_<tool>.log; the summaries kept in the repo are profiles/r2_sanitize_<tool>.txt.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck}
# one small, fast test per kernel family (the full-size tests take minutes under the sanitizer)
TESTS="tests/test_gpu_cost_kl.py::test_cost_kl_golden tests/test_gpu_cost_kl.py::test_cost_kl_packed_teacher_ragged_batched \
tests/test_gpu_smooth_ap.py::test_smooth_ap_golden tests/test_gpu_smooth_ap.py::test_smooth_ap_me_joint_mean_over_the_batch tests/test_gpu_smooth_ap.py::test_smooth_ap_fused_kernel_forced \
tests/test_gpu_depth_rank.py::test_fused_depth_losses_golden tests/test_gpu_depth_rank.py::test_depth_losses_edge_cases tests/test_gpu_sample.py tests/test_gpu_fast_nn.py::test_reciprocal_nn_golden tests/test_gpu_fast_nn.py::test_fast_reciprocal_nns_golden \
tests/test_gpu_teacher_volume.py tests/test_gpu_depth_splat.py tests/test_gpu_eval_argmax.py"
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  timeout 1500 compute-sanitizer --tool "$tool" --error-exitcode 0 --print-limit 20 \
      python -m pytest $TESTS -x -q -p no:cacheprovider > "$log" 2>&1
  echo "== $tool: exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$log" | tail -8
done
