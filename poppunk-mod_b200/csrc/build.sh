#!/bin/bash
# Builds libkdejs.so for sm_100a IN-TREE (the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${KDEJS_OUT:-$HERE/../libkdejs.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -O2 -shared ${KDEJS_PTXAS_V:+-Xptxas -v} ${KDEJS_EXTRA_FLAGS:-} \
    -o "$OUT" "$HERE/kdejs_api.cu" "$HERE/kdejs_io.cpp" -lcudart
echo "built $OUT"
