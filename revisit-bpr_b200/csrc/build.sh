#!/usr/bin/env bash
# Build librbpr.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../librbpr.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=("$here"/api.cu "$here"/train.cu "$here"/score.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xptxas -v -Xcompiler -fPIC,-O3 -shared \
  -o "$out" "${srcs[@]}" 2> "$here/../build.log" || { cat "$here/../build.log"; exit 1; }
echo "built $out"
