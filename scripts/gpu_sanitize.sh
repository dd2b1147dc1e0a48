#!/usr/bin/env bash
# compute-sanitizer over the small-shape GPU parity tests (round-2 to-do: no sanitizer run was made
# in round 1).  Run on the GPU box via:
#   gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh [tag]'
# memcheck: out-of-bounds / misaligned accesses in every kernel the tests launch;
# racecheck: shared-memory hazards (top-k selection, KNN profile, statistics reduction);
# initcheck: reads of uninitialised device memory (scratch buffers the library allocates).
# The full-size tests are deselected (sanitizer slow-down 10-50x); each tool gets its own timeout.
tag="${1:-r2}"
mkdir -p gpurun_out
SEL='not fullsize and not full_size and not experiment'
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
      --log-file "gpurun_out/${tag}_sanitizer_${tool}.log" \
      python -m pytest tests/test_gpu_train.py tests/test_gpu_score.py tests/test_gpu_knn.py \
          tests/test_gpu_adaptive.py tests/test_gpu_score_tc.py -x -q -m gpu -k "$SEL and not (9100 or 8300 or 8200 or 12345 or 8700)" > "gpurun_out/${tag}_sanitizer_${tool}_pytest.log" 2>&1
  echo "$tool: exit $? ; $(grep -c 'ERROR SUMMARY' gpurun_out/${tag}_sanitizer_${tool}.log) summaries ; $(grep 'ERROR SUMMARY' gpurun_out/${tag}_sanitizer_${tool}.log | tail -1)"
  tail -2 "gpurun_out/${tag}_sanitizer_${tool}_pytest.log"
done
