#!/usr/bin/env bash
# round 2, GPU call D (1 GPU): full test suite, full bench line, launch lists, ncu --set full captures
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"
NCU="ncu --clock-control none"
LL="$NCU --metrics gpu__time_duration.sum --csv"
B="python bench.py --no-cpu-baseline --configs none"
timeout 300 $LL -c 300 --log-file gpurun_out/r2d_launches_c2_b65536.csv $B --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2d_launches_c2_b256.csv $B --batch 256 --steps 300 --warmup 50 > /dev/null 2>&1
timeout 400 $LL -c 300 --log-file gpurun_out/r2d_launches_c3_msd_adam.csv $B --shape msd --dim 256 --opt adam --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 300 --log-file gpurun_out/r2d_launches_c4_yelp_adaptive.csv $B --shape yelp --dim 64 --sampler adaptive --steps 8 --warmup 3 > /dev/null 2>&1
timeout 300 $LL -c 200 --log-file gpurun_out/r2d_launches_c5_score.csv python scripts/prof_score.py 2 > /dev/null 2>&1
FULL="$NCU --set full -f"
timeout 400 $FULL -k regex:"bpr_phase_a|bpr_apply" --launch-skip 12 -c 2 -o gpurun_out/r2d_c2_phase_a_apply $B --steps 8 --warmup 3 > /dev/null 2>&1
timeout 500 $FULL -k regex:"bpr_phase_a|bpr_apply" --launch-skip 12 -c 2 -o gpurun_out/r2d_c3_adam_phase_a_apply $B --shape msd --dim 256 --opt adam --steps 8 --warmup 3 > /dev/null 2>&1
timeout 400 $FULL -k regex:"rbpr_sample_adaptive_csr" --launch-skip 6 -c 1 -o gpurun_out/r2d_c4_sample_adaptive $B --shape yelp --dim 64 --sampler adaptive --steps 8 --warmup 3 > /dev/null 2>&1
timeout 400 $FULL -k regex:"score_tc|rescore_rank|select_threshold|build_mask|pack_" --launch-skip 7 -c 7 -o gpurun_out/r2d_c5_score python scripts/prof_score.py 2 > /dev/null 2>&1
timeout 400 $FULL -k regex:"bpr_small_steps" --launch-skip 3 -c 1 -o gpurun_out/r2d_c2_small_steps $B --batch 256 --steps 600 --warmup 100 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
python - <<P
import json
d=json.load(open("gpurun_out/r2d_bench.json")); c=d["configs"]
print("headline %.4g  e2e %.4g  ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
for k,v in c.items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","gpu_launches","error") if x in v}, (v.get("roofline") or {}).get("frac"))
P
