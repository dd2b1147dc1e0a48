#!/usr/bin/env bash
# round 2, GPU call F (1 GPU): full test suite, bench, launch lists after the kernel rework
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -25 gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"
NCU="ncu --clock-control none"
LL="$NCU --metrics gpu__time_duration.sum --csv"
B="python bench.py --no-cpu-baseline --configs none"
timeout 300 $LL -c 300 --log-file gpurun_out/r2f_launches_c2_b65536.csv $B --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2f_launches_c2_b256.csv $B --batch 256 --steps 300 --warmup 50 > /dev/null 2>&1
timeout 300 $LL -c 300 --log-file gpurun_out/r2f_launches_c4_yelp_adaptive.csv $B --shape yelp --dim 64 --sampler adaptive --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2f_launches_c5_score.csv python scripts/prof_score.py 2 > /dev/null 2>&1
for f in c2_b65536 c2_b256 c4_yelp_adaptive c5_score; do echo "=== $f"; python scripts/launch_summary.py gpurun_out/r2f_launches_$f.csv | grep -v "native::\|at::\|at_cuda\|CUB_\|randperm\|elementwise"; done
python - <<P
import json
d=json.load(open("gpurun_out/r2f_bench.json")); c=d["configs"]
print("headline %.4g  e2e %.4g  ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error") if x in v}, (v.get("roofline") or {}).get("frac"))
P
