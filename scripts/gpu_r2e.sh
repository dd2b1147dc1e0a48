#!/usr/bin/env bash
# round 2, GPU call E (2 GPUs): fused exchange with folded barriers — parity, DDP experiment (with hang dumps), bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== parity, fused exchange"; RBPR_FUSED_EXCHANGE=1 timeout 200 $TR --master-port 29542 tests/tools/check_multi_gpu.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
echo "== experiment ddp"; RBPR_HANG_DUMP_S=60 timeout 300 $TR --master-port 29544 tests/tools/check_experiment_ddp.py > gpurun_out/r2e_ddp.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2e_ddp.log | tail -40
echo "== bench N=2 fused"
timeout 400 $TR --master-port 29551 bench.py --gpus 2 --configs c3,c5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "exit $?"
python - <<P
import json
for ln in open("gpurun_out/r2e_bench_n2.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("value %.4g ms/step %.4f launches %d allreduces %d fused %d parity %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["nccl_allreduces"], d["fused_exchanges"], d.get("parity_check")))
        for k,v in d["configs"].items(): print(k, {x: v.get(x) for x in ("value","ms_per_step","ms","error","fused_exchanges") if x in v})
P
tail -5 gpurun_out/r2e_bench_n2.err
