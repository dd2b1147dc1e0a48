#!/usr/bin/env bash
# Build librbpr.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${RBPR_OUT:-$here/../librbpr.so}"
obj="${RBPR_OBJ:-$here/../build}"
mkdir -p "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(${RBPR_DEFS:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -Xcompiler -fPIC,-O3)
pids=()
for f in api train train_small train_sgd train_adam train_sgdm train_rms score score_tc dropin knn adaptive comm exchange ingest; do
  ( "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$obj/$f.o" > "$obj/$f.log" 2>&1 || { cat "$obj/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out" "$obj"/{api,train,train_small,train_sgd,train_adam,train_sgdm,train_rms,score,score_tc,dropin,knn,adaptive,comm,exchange,ingest}.o -ldl -Xlinker -z -Xlinker defs
cat "$obj"/*.log > "$here/../build.log"
echo "built $out"
