#!/usr/bin/env bash
# round 2, GPU call U (8-GPU box, one short run): multicast pull + UNICAST push at N=8 (compare with z_bench_n8_full: 94.4 us/step)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RBPR_FX_MC_PUSH=0 RBPR_FX_TRACE=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --configs none > gpurun_out/r2u_bench_n8_ucpush.json 2> gpurun_out/r2u_bench_n8_ucpush.err
python - <<P
import json
for ln in open("gpurun_out/r2u_bench_n8_ucpush.json"):
    if ln.startswith("{"):
        d=json.loads(ln)
        print("ucpush ms/step %.4f value %.4g e2e %.4g fused %d [%s] parity %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["fused_exchanges"], d.get("exchange"), d.get("parity_check")))
P
grep -o "\[rbpr fx trace\] rank [0-9]: 60 exchanges, median[^\[]*next exchange start (phase A etc.) [0-9.]*" gpurun_out/r2u_bench_n8_ucpush.err | awk 'length($0) < 330' | tail -3
tail -c 300 gpurun_out/r2u_bench_n8_ucpush.err
