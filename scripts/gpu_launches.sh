#!/usr/bin/env bash
# usage: bash scripts/gpu_launches.sh tag [bench args]   -> gpurun_out/<tag>_launches.csv
tag="${1:-launches}"; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" \
    > gpurun_out/${tag}_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches.csv
