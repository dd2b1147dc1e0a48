#!/usr/bin/env bash
# round 2, GPU call P (N GPUs, default 2): exchange kernel v2 (element-granular reduce, user half fused) —
# parity, per-phase trace of the exchange kernel, bench fused / split-users / NCCL
N="${1:-2}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity, symmetric buffer + multicast"; RBPR_FX_TRACE=1 timeout 200 $TR --master-port 29542 tests/tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | grep -v "fx trace" | tail -6
echo "== parity, symmetric buffer, unicast"; RBPR_FX_MULTICAST=0 timeout 200 $TR --master-port 29543 tests/tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
echo "== parity, cudaIpc binding"; RBPR_FX_SYMM=0 timeout 200 $TR --master-port 29545 tests/tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
echo "== experiment ddp"; RBPR_HANG_DUMP_S=60 timeout 300 $TR --master-port 29544 tests/tools/check_experiment_ddp.py > gpurun_out/r2p_ddp.log 2>&1; echo "exit $?"; grep "experiment ddp\|Timeout\|Error" gpurun_out/r2p_ddp.log | head -12
run() {  # $1 = tag, rest = env
  local tag=$1; shift 1
  env "$@" timeout 420 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N > gpurun_out/r2p_bench_n${N}_$tag.json 2> gpurun_out/r2p_bench_n${N}_$tag.err
  echo "bench $tag exit $?"
  python - <<P
import json
ok=False
for ln in open("gpurun_out/r2p_bench_n${N}_$tag.json"):
    if ln.startswith("{"):
        d=json.loads(ln); ok=True
        print("$tag: value %.4g ms/step %.4f e2e %.4g launches %d allreduces %d fused %d [%s] parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["nccl_allreduces"], d["fused_exchanges"], d.get("exchange"), d.get("parity_check")))
        for k,v in d["configs"].items(): print("   ", k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","fused_exchanges","nccl_allreduces") if x in v})
if not ok: print(open("gpurun_out/r2p_bench_n${N}_$tag.err").read()[-1500:])
P
  grep "fx trace" gpurun_out/r2p_bench_n${N}_$tag.err | grep "rank 0" | tail -4
}
run mc_trace RBPR_FX_TRACE=1 RBPR_BENCH_CONFIGS=none
run uc_trace RBPR_FX_TRACE=1 RBPR_FX_MULTICAST=0 RBPR_BENCH_CONFIGS=none
run mc
run ipc RBPR_FX_SYMM=0 RBPR_BENCH_CONFIGS=none
