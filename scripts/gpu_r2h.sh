#!/usr/bin/env bash
# round 2, GPU call H (1 GPU): tests, bench, scoring launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -30 gpurun_out/r2h_pytest.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
NCU="ncu --clock-control none"
LL="$NCU --metrics gpu__time_duration.sum --csv"
timeout 300 $LL -c 200 --log-file gpurun_out/r2h_launches_c5_score.csv python scripts/prof_score.py 2 > gpurun_out/r2h_prof_score.log 2>&1
tail -1 gpurun_out/r2h_prof_score.log
python scripts/launch_summary.py gpurun_out/r2h_launches_c5_score.csv | grep -v "native::\|at::\|at_cuda\|CUB_\|randperm\|elementwise"
python - <<P
import json
d=json.load(open("gpurun_out/r2h_bench.json")); c=d["configs"]
print("headline %.4g  e2e %.4g  ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error") if x in v}, (v.get("roofline") or {}).get("frac"))
print(json.dumps(c.get("e2e_experiment"))[:900])
P
