#!/usr/bin/env bash
# round 2, GPU call S (N GPUs): chunk-major multicast pipeline on/off, defaults
N="${1:-2}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity (multicast forced, chunked)"; RBPR_FX_MULTICAST=1 timeout 200 $TR --master-port 29542 tests/tools/check_multi_gpu.py 2>&1 | grep "multi-gpu check\|Error\|error" | tail -6
run() {  # $1 = tag, $2 = extra bench args, rest = env
  local tag=$1 extra=$2; shift 2
  env "$@" timeout 300 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N $extra > gpurun_out/r2s_bench_n${N}_$tag.json 2> gpurun_out/r2s_bench_n${N}_$tag.err
  python - <<P
import json
ok=False
for ln in open("gpurun_out/r2s_bench_n${N}_$tag.json"):
    if ln.startswith("{"):
        d=json.loads(ln); ok=True
        print("%-14s ms/step %.4f value %.4g e2e %.4g launches %d fused %d [%s] parity %s" % ("$tag", d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["fused_exchanges"], d.get("exchange"), d.get("parity_check")))
        for k,v in d["configs"].items(): print("   ", k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","exchange") if x in v})
if not ok: print("$tag FAILED", open("gpurun_out/r2s_bench_n${N}_$tag.err").read()[-1200:])
P
  grep -o "\[rbpr fx trace\] rank 0: 60 exchanges[^\[]*" gpurun_out/r2s_bench_n${N}_$tag.err | tail -1 | cut -c1-330
}
Q="--configs none --no-parity-check"
run mc_c4 "--configs none" RBPR_FX_TRACE=1 RBPR_FX_MULTICAST=1
run mc_c1 "$Q" RBPR_FX_TRACE=1 RBPR_FX_MULTICAST=1 RBPR_FX_MC_CHUNKS=1
if [ "$N" = "2" ]; then run default "$Q"; else run full ""; fi
