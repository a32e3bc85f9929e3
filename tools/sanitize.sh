#!/bin/bash
# compute-sanitizer over the small parity cases (the reference has no race detection; this is
# ours).  Meant for the GPU box, one tool per call:
#   gpurun --timeout 900 -- 'tools/sanitize.sh memcheck  > gpurun_out/sanitize_memcheck.log 2>&1'
#   gpurun --timeout 900 -- 'tools/sanitize.sh racecheck > gpurun_out/sanitize_racecheck.log 2>&1'
# The selection keeps every kernel family on the path (tensor-core pair at several ranks, general kernel,
# grid walk / grouped walk / table walk, PI evaluation, off-grid entries, fused multi-destination stores)
# at sizes the sanitizer finishes in minutes; the full-size and solver tests are left out.
set -u
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
extra=()
case "$tool" in
  memcheck)  extra=(--leak-check no --report-api-errors no) ;;
  racecheck) extra=(--racecheck-report analysis) ;;
  synccheck|initcheck) ;;
  *) echo "usage: $0 memcheck|racecheck|synccheck|initcheck" >&2; exit 2 ;;
esac
exec compute-sanitizer --tool "$tool" "${extra[@]}" --error-exitcode 97 --target-processes all \
  python -m pytest -q -x -p no:cacheprovider -m gpu \
    "tests/test_gpu_parity.py::test_vi_debug_against_oracle" \
    "tests/test_gpu_parity.py::test_pi_two_subiterations" \
    "tests/test_gpu_paths.py::test_tensor_core_path_rank_sweep" \
    "tests/test_gpu_paths.py::test_mixed_ranks_and_ragged_groups" \
    "tests/test_gpu_paths.py::test_general_kernel_large_ranks" \
    "tests/test_gpu_paths.py::test_grid_walk_equals_table_walk" \
    "tests/test_gpu_paths.py::test_fused_gather_stores" \
    "tests/test_policy_eval.py::test_policy_eval_gpu_equals_oracle"
