#!/usr/bin/env bash
# round 2, GPU call N8 (8 GPUs): parity + bench at N=8 (fused exchange and NCCL fallback), N=4, DDP experiment
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # $1 = ranks, $2 = tag, rest = env
  local n=$1 tag=$2; shift 2
  env "$@" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n > gpurun_out/r2n_bench_$tag.json 2> gpurun_out/r2n_bench_$tag.err
  echo "bench $tag exit $?"
  python - <<P
import json
ok=False
for ln in open("gpurun_out/r2n_bench_$tag.json"):
    if ln.startswith("{"):
        d=json.loads(ln); ok=True
        print("$tag: value %.4g ms/step %.4f e2e %.4g launches %d allreduces %d fused %d parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["nccl_allreduces"], d["fused_exchanges"], d.get("parity_check")))
        for k,v in d["configs"].items(): print("   ", k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","fused_exchanges","nccl_allreduces") if x in v})
if not ok: print(open("gpurun_out/r2n_bench_$tag.err").read()[-1500:])
P
}
run 8 n8_fused RBPR_FUSED_EXCHANGE=1
run 8 n8_nccl RBPR_FUSED_EXCHANGE=0 RBPR_BENCH_CONFIGS=none
run 4 n4_fused RBPR_FUSED_EXCHANGE=1 RBPR_BENCH_CONFIGS=none
echo "== experiment ddp (8 ranks)"; RBPR_HANG_DUMP_S=90 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tests/tools/check_experiment_ddp.py > gpurun_out/r2n_ddp8.log 2>&1; echo "exit $?"; grep "experiment ddp\|Timeout\|Error" gpurun_out/r2n_ddp8.log | head -12
