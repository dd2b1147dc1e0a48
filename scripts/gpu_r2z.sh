#!/usr/bin/env bash
# round 2, final 1-GPU call: full GPU test suite, both bench arms, launch lists of every config (profiles/round2/z_*)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2z_pytest.log
tail -6 gpurun_out/r2z_pytest.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "reference arm exit $?"; tail -c 600 gpurun_out/r2z_bench_reference.json
NCU="ncu --clock-control none"
LL="$NCU --metrics gpu__time_duration.sum --csv"
B="python bench.py --no-cpu-baseline --configs none"
timeout 300 $LL -c 300 --log-file gpurun_out/r2z_launches_c2_b65536.csv $B --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2z_launches_c2_b256.csv $B --batch 256 --steps 300 --warmup 50 > /dev/null 2>&1
timeout 300 $LL -c 300 --log-file gpurun_out/r2z_launches_c3_msd_adam.csv $B --shape msd --dim 256 --opt adam --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 300 --log-file gpurun_out/r2z_launches_c4_yelp_adaptive.csv $B --shape yelp --dim 64 --sampler adaptive --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2z_launches_c5_score.csv python scripts/prof_score.py 2 > /dev/null 2>&1
for f in c2_b65536 c2_b256 c3_msd_adam c4_yelp_adaptive c5_score; do echo "=== $f"; python scripts/launch_summary.py gpurun_out/r2z_launches_$f.csv | grep -v "native::\|at::\|at_cuda\|CUB_\|randperm\|elementwise" | tee gpurun_out/r2z_launches_$f.txt; done
python - <<P
import json
d=json.load(open("gpurun_out/r2z_bench.json")); c=d["configs"]
print("headline %.4g  e2e %.4g  ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error") if x in v}, (v.get("roofline") or {}).get("frac"))
P
