#!/usr/bin/env bash
# round 2, GPU call G (1 GPU): tests, bench, small-batch cluster A/B, ncu of the scoring kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -30 gpurun_out/r2g_pytest.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $?"
for cl in 8 16; do RBPR_SMALL_CLUSTER=$cl timeout 200 python bench.py --configs c2_b256 --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['c2_b256']; print('cluster $cl: b256 %.4g triples/s %.3f us/step' % (c['value'], 1e3*c['ms_per_step']))"; done
NCU="ncu --clock-control none"
LL="$NCU --metrics gpu__time_duration.sum --csv"
B="python bench.py --no-cpu-baseline --configs none"
timeout 300 $LL -c 300 --log-file gpurun_out/r2g_launches_c2_b65536.csv $B --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 300 --log-file gpurun_out/r2g_launches_c4_yelp_adaptive.csv $B --shape yelp --dim 64 --sampler adaptive --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2g_launches_c5_score.csv python scripts/prof_score.py 2 > gpurun_out/r2g_prof_score.log 2>&1
tail -1 gpurun_out/r2g_prof_score.log
for f in c2_b65536 c4_yelp_adaptive c5_score; do echo "=== $f"; python scripts/launch_summary.py gpurun_out/r2g_launches_$f.csv | grep -v "native::\|at::\|at_cuda\|CUB_\|randperm\|elementwise"; done
timeout 400 $NCU --set full -f -k regex:"score_tc|rescore_rank|select_threshold" --launch-skip 3 -c 4 -o gpurun_out/r2g_c5_score python scripts/prof_score.py 2 > /dev/null 2>&1
python - <<P
import json
d=json.load(open("gpurun_out/r2g_bench.json")); c=d["configs"]
print("headline %.4g  e2e %.4g  ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error") if x in v}, (v.get("roofline") or {}).get("frac"))
P
