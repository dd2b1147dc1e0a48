#!/usr/bin/env bash
# compute-sanitizer over small-shape GPU parity tests.  Run on the GPU box via:
#   gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh [tag]'
# memcheck: out-of-bounds / misaligned accesses in every kernel the tests launch;
# racecheck: shared-memory hazards (top-k selection, statistics reduction, KNN profile).
# Full-size tests are left out (sanitizer slow-down 10-50x); every group has its own timeout, and a
# group cut off by it says so instead of reporting a summary.
tag="${1:-r2}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
grp() {  # $1 tool, $2 name, $3 timeout, $4 -k expression, rest: test files
  local tool=$1 name=$2 to=$3 sel=$4; shift 4
  local log="gpurun_out/${tag}_sanitizer_${tool}_${name}.log"
  timeout "$to" compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 10 --log-file "$log" \
      python -m pytest "$@" -x -q -m gpu -k "$sel" > "gpurun_out/${tag}_sanitizer_${tool}_${name}_pytest.log" 2>&1
  local rc=$?
  echo "$tool/$name: exit $rc$([ $rc -eq 124 ] && echo ' (cut off by the timeout)') ; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' "$log" 2>/dev/null | tail -1) ; pytest: $(tail -1 gpurun_out/${tag}_sanitizer_${tool}_${name}_pytest.log)"
}
ALL="tests/test_gpu_train.py tests/test_gpu_score.py tests/test_gpu_score_tc.py tests/test_gpu_adaptive.py tests/test_gpu_knn.py tests/test_gpu_dropin.py tests/test_gpu_experiment.py"
SMALL="not fullsize and not full_size"
grp memcheck all 200 "$SMALL" $ALL
grp racecheck all 150 "$SMALL" $ALL
grp synccheck all 100 "$SMALL" $ALL
grp initcheck train_score 100 "$SMALL" tests/test_gpu_train.py tests/test_gpu_score.py tests/test_gpu_score_tc.py
