#!/usr/bin/env bash
# round 2, GPU call B (2 GPUs): data-parallel parity (NCCL path and fused peer-memory exchange),
# symmetric-memory / multicast probe, short 2-GPU bench of both exchange variants
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== parity, NCCL exchange" ; RBPR_FUSED_EXCHANGE=0 timeout 300 $TR --master-port 29541 tests/tools/check_multi_gpu.py 2>&1 | tail -4
echo "== parity, fused exchange"; RBPR_FUSED_EXCHANGE=1 timeout 300 $TR --master-port 29542 tests/tools/check_multi_gpu.py 2>&1 | tail -6
echo "== experiment ddp"; timeout 400 $TR --master-port 29544 tests/tools/check_experiment_ddp.py 2>&1 | tail -8
echo "== probe"; timeout 200 $TR --master-port 29543 scripts/probe_symm.py 2>&1 | tail -12
for fx in 0 1; do
  echo "== bench N=2 fused=$fx"
  RBPR_FUSED_EXCHANGE=$fx timeout 400 $TR --master-port 2955$fx bench.py --gpus 2 --configs none --no-parity-check > gpurun_out/r2b_bench_n2_fx$fx.json 2> gpurun_out/r2b_bench_n2_fx$fx.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_n2_fx$fx.json")); print("value %.4g ms/step %.4f launches %d allreduces %d fused %d" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["nccl_allreduces"], d["fused_exchanges"]))
except Exception as e: print("no json", e); print(open("gpurun_out/r2b_bench_n2_fx$fx.err").read()[-1500:])
P
done
