#!/usr/bin/env bash
# round 2, GPU call T (8-GPU box): final defaults at N=8 (full line), exchange trace (medians), experimental chunked pipeline
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # $1 = ranks, $2 = tag, $3 = extra bench args, rest = env
  local n=$1 tag=$2 extra=$3; shift 3
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n $extra > gpurun_out/r2t_bench_$tag.json 2> gpurun_out/r2t_bench_$tag.err
  python - <<P
import json
ok=False
for ln in open("gpurun_out/r2t_bench_$tag.json"):
    if ln.startswith("{"):
        d=json.loads(ln); ok=True
        print("%-14s ms/step %.4f value %.4g e2e %.4g launches %d fused %d nccl %d [%s] parity %s" % ("$tag", d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["fused_exchanges"], d["nccl_allreduces"], d.get("exchange"), d.get("parity_check")))
        for k,v in d["configs"].items(): print("   ", k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","exchange") if x in v})
if not ok: print("$tag FAILED", open("gpurun_out/r2t_bench_$tag.err").read()[-1200:])
P
  grep -o "\[rbpr fx trace\] rank [0-9]: 60 exchanges, median[^\[]*next exchange start (phase A etc.) [0-9.]*" gpurun_out/r2t_bench_$tag.err | awk 'length($0) < 330' | tail -3
}
Q="--configs none --no-parity-check"
run 8 n8_full ""
run 8 n8_trace "$Q" RBPR_FX_TRACE=1
run 8 n8_chunks4 "$Q" RBPR_FX_MC_CHUNKS=4
run 8 n8_steps240 "$Q --steps 240"
