#!/usr/bin/env bash
# round 2, GPU call A: test suite + default bench (all configs) on one B200
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2a_bench.err
head -c 3000 gpurun_out/r2a_bench.json
