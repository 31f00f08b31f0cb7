#!/usr/bin/env bash
# Builds experimental variants of the library into ab_libs/<name>.so (git-ignored; they travel to the GPU box with gpurun).
# A variant = the in-tree sources + one -D flag; the default build does not contain any of them (tools/build_lib.sh).
#   tools/gpu_ab.sh then runs the parity tests and the bench for every library under ab_libs/ and for the in-tree one.
# Usage: bash tools/build_variants.sh [FLAG ...]      (default: every flag listed below)
#   RN_EXP_GEMM_4X12    tile_gemm with 4 rows x 12 columns per thread on 128 threads (fewer shared-memory wavefronts per FMA)
#   RN_EXP_GEMM_FFMA2   tile_gemm with packed FMAs (fma.rn.f32x2): half the FMA instructions, bit-identical results
#   RN_EXP_RANGESUM16   crown head sums with 16 rows in flight per trip (large crowns: C3's root adds 480 heads)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/rapidnet_b200/csrc"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS_ALL=()
if [ $# -gt 0 ]; then FLAGS_ALL=("$@"); fi
mkdir -p "$ROOT/ab_libs"
for f in "${FLAGS_ALL[@]}"; do
  out="$ROOT/ab_libs/$(echo "$f" | tr 'A-Z' 'a-z' | sed 's/^rn_exp_//').so"
  "$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC --shared \
      -D"$f" -I"$ROOT/include" -o "$out" "$SRC/rn_api.cu" "$SRC/rn_factor.cu" "$SRC/rn_affine.cu" "$SRC/rn_apg.cu" "$SRC/rn_batched.cu" "$SRC/rn_persist.cu" \
      -lcusolver -lcudart
  echo "built $out"
done
