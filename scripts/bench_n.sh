#!/usr/bin/env bash
# usage: bash scripts/bench_n.sh tag N [bench args]  -> gpurun_out/<tag>_n<N>.json (one JSON line)
tag="$1"; n="$2"; shift; shift
mkdir -p gpurun_out
if [ "$n" = "1" ]; then
  timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/${tag}_n${n}.json 2> gpurun_out/${tag}_n${n}.err
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n "$@" > gpurun_out/${tag}_n${n}.json 2> gpurun_out/${tag}_n${n}.err
fi
tail -3 gpurun_out/${tag}_n${n}.err | cut -c1-300
python - "$tag" "$n" <<'PY'
import json, sys
tag, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/{tag}_n{n}.json").read().strip().splitlines()[-1])
    print(f"N={d['n_gpus']} B={d['config']['batch']} value={d['value']/1e6:.1f}M us/step={d['ms_per_step']*1e3:.1f} "
          f"kernel_us={d['roofline']['kernel_ms_avg']*1e3:.1f} e2e={d['e2e']['value']/1e6:.1f}M launches={d['gpu_launches']} "
          f"allreduces={d.get('nccl_allreduces')} clocks={d['clocks']}")
except Exception as e:
    print("no result:", e)
PY
