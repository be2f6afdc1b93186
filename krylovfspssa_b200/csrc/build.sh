#!/bin/bash
# Builds krylovfspssa_b200/libkfsp.so for sm_100a (the only target).  nvcc cross-compiles
# without a GPU.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${KFSP_OUT:-$HERE/../libkfsp.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
NCCL_FLAGS=()
# NCCL (torch-bundled wheel): needed for the multi-GPU path only; found automatically when importable
if [ -z "${KFSP_NCCL_INC:-}" ] && command -v python >/dev/null 2>&1; then
  NCCL_BASE="$(python -c 'import nvidia.nccl as n, os; print(list(n.__path__)[0])' 2>/dev/null || true)"
  if [ -n "$NCCL_BASE" ] && [ -f "$NCCL_BASE/include/nccl.h" ]; then
    export KFSP_NCCL_INC="$NCCL_BASE/include" KFSP_NCCL_LIB="$NCCL_BASE/lib"
  fi
fi
if [ -n "${KFSP_NCCL_INC:-}" ] && [ -n "${KFSP_NCCL_LIB:-}" ]; then
  NCCL_FLAGS=(-DKFSP_WITH_NCCL "-I$KFSP_NCCL_INC" "-L$KFSP_NCCL_LIB" -l:libnccl.so.2 "-Xlinker" "-rpath=$KFSP_NCCL_LIB")
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared \
  -o "$OUT" "$HERE/kfsp.cu" "$HERE/model_host.cpp" "${NCCL_FLAGS[@]}" "$@"
echo "built $OUT"
