#!/usr/bin/env bash
# round 2, GPU call R (8-GPU box): which exchange binding / role split is fastest at N=8 and N=4, then
# the full bench line (c3, c5, parity) at N=8 with the winner, and the DDP experiment at 8 ranks.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {  # $1 = ranks, $2 = tag, $3 = extra bench args, rest = env
  local n=$1 tag=$2 extra=$3; shift 3
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $n $extra > gpurun_out/r2r_bench_$tag.json 2> gpurun_out/r2r_bench_$tag.err
  python - <<P
import json
ok=False
for ln in open("gpurun_out/r2r_bench_$tag.json"):
    if ln.startswith("{"):
        d=json.loads(ln); ok=True
        print("%-14s ms/step %.4f value %.4g e2e %.4g launches %d fused %d nccl %d [%s] parity %s" % ("$tag", d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["fused_exchanges"], d["nccl_allreduces"], d.get("exchange"), d.get("parity_check")))
        for k,v in d["configs"].items(): print("   ", k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","exchange") if x in v})
        open("gpurun_out/r2r_times.txt","a").write("$tag %.6f\n" % d["ms_per_step"])
if not ok: print("$tag FAILED", open("gpurun_out/r2r_bench_$tag.err").read()[-1200:])
P
  grep "fx trace" gpurun_out/r2r_bench_$tag.err | grep "rank 0" | tail -1 | cut -c1-330
}
rm -f gpurun_out/r2r_times.txt
Q="--configs none --no-parity-check"
run 8 n8_mc_s2 "--configs none" RBPR_FX_TRACE=1
run 8 n8_uc_s2 "--configs none" RBPR_FX_TRACE=1 RBPR_FX_MULTICAST=0
run 8 n8_mc_s4 "$Q" RBPR_FX_XCHG_CTAS=4
run 4 n4_mc_s2 "$Q"
run 4 n4_uc_s2 "$Q" RBPR_FX_MULTICAST=0
best=$(grep "^n8_" gpurun_out/r2r_times.txt | sort -k2 -n | head -1 | cut -d' ' -f1)
echo "== fastest at N=8: $best"
case "$best" in
  n8_uc_s2) run 8 n8_full "" RBPR_FX_MULTICAST=0 ;;
  n8_mc_s4) run 8 n8_full "" RBPR_FX_XCHG_CTAS=4 ;;
  *) run 8 n8_full "" ;;
esac
echo "== experiment ddp (8 ranks)"; RBPR_HANG_DUMP_S=90 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tests/tools/check_experiment_ddp.py > gpurun_out/r2r_ddp8.log 2>&1; echo "exit $?"; grep "experiment ddp\|Timeout\|Error" gpurun_out/r2r_ddp8.log | head -12
