#!/usr/bin/env bash
# Builds rapidnet_b200/librapidnet_b200.so (the C-ABI library, CUDA sm_100a) in-tree.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/rapidnet_b200/csrc"
OUT="$ROOT/rapidnet_b200/librapidnet_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX="${HOSTCXX:-/usr/bin/g++}"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin "$HOSTCXX"
       -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v --shared
       -I"$ROOT/include")
"$NVCC" "${FLAGS[@]}" -o "$OUT" "$SRC/rn_api.cu" "$SRC/rn_factor.cu" "$SRC/rn_affine.cu" "$SRC/rn_apg.cu" "$SRC/rn_batched.cu" "$SRC/rn_persist.cu" \
    -lcusolver -lcudart "$@"
echo "built $OUT"
