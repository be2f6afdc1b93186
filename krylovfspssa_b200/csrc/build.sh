#!/bin/bash
# Builds krylovfspssa_b200/libkfsp.so for sm_100a (the only target).  nvcc cross-compiles
# without a GPU.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libkfsp.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
NCCL_FLAGS=()
if [ -n "${KFSP_NCCL_INC:-}" ] && [ -n "${KFSP_NCCL_LIB:-}" ]; then
  NCCL_FLAGS=(-DKFSP_WITH_NCCL "-I$KFSP_NCCL_INC" "-L$KFSP_NCCL_LIB" -l:libnccl.so.2 "-Xlinker" "-rpath=$KFSP_NCCL_LIB")
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared \
  -o "$OUT" "$HERE/kfsp.cu" "$HERE/model_host.cpp" "${NCCL_FLAGS[@]}" "$@"
echo "built $OUT"
