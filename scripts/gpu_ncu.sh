#!/usr/bin/env bash
# usage: bash scripts/gpu_ncu.sh tag [kernel-regex] [extra bench args]
tag="${1:-prof}"; kre="${2:-bpr_phase_a}"; shift; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${kre} -s 4 -c 1 \
    -o gpurun_out/${tag} -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" \
    > gpurun_out/${tag}.log 2>&1
tail -2 gpurun_out/${tag}.log | cut -c1-300
