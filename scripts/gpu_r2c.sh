#!/usr/bin/env bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# 1. everything except the new tensor scoring path
RBPR_NO_TC_SCORE=1 timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_score_tc.py > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
# 2. the tensor scoring path, in its own process
timeout 600 python -m pytest tests/test_gpu_score_tc.py tests/test_gpu_score.py "tests/test_gpu_fullsize.py::test_config5_ml20m_full_catalog_scoring_parity" -q -x > gpurun_out/r2c_pytest_tc.log 2>&1; echo "pytest-tc exit $?" >> gpurun_out/r2c_pytest_tc.log
tail -25 gpurun_out/r2c_pytest_tc.log
timeout 300 python bench.py --configs c2_b256,c5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"
python - <<P
import json
d=json.load(open("gpurun_out/r2c_bench.json")); c=d["configs"]
print("headline %.4g" % d["value"])
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error","ndcg@100","effective_tflops")})
P
