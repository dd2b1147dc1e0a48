#!/usr/bin/env bash
# Run on the GPU box via: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
# tests -> smoke -> bench (both arms) -> ncu launch list -> ncu full capture of the dominant kernel
tag="${1:-r01}"
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 10 --warmup 3 2>gpurun_out/${tag}_ref.err | tee gpurun_out/${tag}_bench_reference.json | cut -c1-400
python bench.py 2>gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | cut -c1-600
tail -5 gpurun_out/${tag}_bench.err
if [ -z "$SKIP_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bpr_phase_a -s 4 -c 1 \
    -o gpurun_out/${tag}_phase_a -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
fi
