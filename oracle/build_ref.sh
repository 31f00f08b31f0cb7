#!/usr/bin/env bash
# Builds the reference's own CUDA/cuBLAS implementation from the sources where they lie
# (/root/reference/src, read-only) into oracle/_ref/ (git-ignored, travels to the GPU box):
#   oracle/_ref/ref_driver   -- oracle/ref_driver.cu + the reference's Engine/SmpcController/loaders
#   oracle/_ref/ref_loop     -- oracle/ref_loop.cu: the reference's closed loop (src/main.cu:27-63) with controls, states, KPIs dumped
#   oracle/_ref/ref_tests    -- the reference's own main() with TESTING=1 (src/main.cu + src/test/*.cu)
# No reference source is copied or modified; the one toolchain incompatibility (abs(size_t)) is
# handled by force-including oracle/ref_compat.h.  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RAPIDNET_REFERENCE:-/root/reference}/src"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ ! -d "$REF" ]; then echo "build_ref: $REF not present (GPU box: use the prebuilt oracle/_ref)"; exit 0; fi
mkdir -p "$OUT/obj"
FLAGS=(-std=c++14 -O2 -w -gencode arch=compute_100,code=sm_100 -I"$REF" -include "$HERE/ref_compat.h")
for f in DwnNetwork ScenarioTree Forecaster SmpcConfiguration Engine Utilities SmpcController; do
    if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/$f.cu" -nt "$OUT/obj/$f.o" ]; then
        "$NVCC" "${FLAGS[@]}" -c "$REF/$f.cu" -o "$OUT/obj/$f.o"
    fi
done
CORE=("$OUT/obj/DwnNetwork.o" "$OUT/obj/ScenarioTree.o" "$OUT/obj/Forecaster.o" "$OUT/obj/SmpcConfiguration.o"
      "$OUT/obj/Engine.o" "$OUT/obj/Utilities.o" "$OUT/obj/SmpcController.o")
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_driver.cu" -o "$OUT/obj/ref_driver.o"
"$NVCC" -gencode arch=compute_100,code=sm_100 "${CORE[@]}" "$OUT/obj/ref_driver.o" -o "$OUT/ref_driver" -lcublas -lcusolver
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_loop.cu" -o "$OUT/obj/ref_loop.o"
"$NVCC" -gencode arch=compute_100,code=sm_100 "${CORE[@]}" "$OUT/obj/ref_loop.o" -o "$OUT/ref_loop" -lcublas -lcusolver
for f in Testing TestSmpcController; do
    "$NVCC" "${FLAGS[@]}" -c "$REF/test/$f.cu" -o "$OUT/obj/$f.o"
done
"$NVCC" "${FLAGS[@]}" -c "$REF/main.cu" -o "$OUT/obj/main.o"
"$NVCC" -gencode arch=compute_100,code=sm_100 "${CORE[@]}" "$OUT/obj/Testing.o" "$OUT/obj/TestSmpcController.o" "$OUT/obj/main.o" \
    -o "$OUT/ref_tests" -lcublas -lcusolver
echo "built $OUT/ref_driver $OUT/ref_loop $OUT/ref_tests"
