#!/usr/bin/env bash
# Drop-in proof: the reference's OWN test sources (src/test/Testing.cu, src/test/TestSmpcController.cu), unchanged and compiled
# from where they lie, against rapidnet-b200's class surface (rapidnet_b200/host/shim/*.cuh = the reference's header names over
# rapidnet_host.hpp) instead of the reference's classes.  The sources include "../SmpcController.cuh" relative to their own
# directory, so a build tree of SYMLINKS puts the shim headers where the reference's headers would be:
#     _ref/shim_src/test/{Testing,TestSmpcController}.{cu,cuh} -> /root/reference/src/test/...
#     _ref/shim_src/*.cuh, Configuration.h                      -> rapidnet_b200/host/shim/...
#     _ref/shim_src/rapidjson                                   -> the reference's vendored rapidjson headers (third party)
# oracle/shim_main.cu runs the in-scope part of the reference's main() (src/main.cu:14-19: loaders, Engine, APG steps).
# Output: oracle/_ref/ref_shim_tests (git-ignored, travels to the GPU box).  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${RAPIDNET_REFERENCE:-/root/reference}/src"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ ! -d "$REF" ]; then echo "build_shim_tests: $REF not present (GPU box: use the prebuilt oracle/_ref)"; exit 0; fi
T="$OUT/shim_src"
rm -rf "$T"; mkdir -p "$T/test" "$OUT/obj"
for f in Testing.cu Testing.cuh TestSmpcController.cu TestSmpcController.cuh; do ln -s "$REF/test/$f" "$T/test/$f"; done
for f in Configuration.h DwnNetwork.cuh ScenarioTree.cuh Forecaster.cuh SmpcConfiguration.cuh Engine.cuh SmpcController.cuh Utilities.cuh; do
    ln -s "$ROOT/rapidnet_b200/host/shim/$f" "$T/$f"
done
ln -s "$REF/rapidjson" "$T/rapidjson"
FLAGS=(-std=c++17 -O2 -w -gencode arch=compute_100a,code=sm_100a -I"$T" -I"$ROOT/rapidnet_b200/host" -I"$ROOT/include")
for f in Testing TestSmpcController; do
    "$NVCC" "${FLAGS[@]}" -c "$T/test/$f.cu" -o "$OUT/obj/shim_$f.o"
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/shim_main.cu" -o "$OUT/obj/shim_main.o"
"$NVCC" -gencode arch=compute_100a,code=sm_100a "$OUT/obj/shim_Testing.o" "$OUT/obj/shim_TestSmpcController.o" "$OUT/obj/shim_main.o" \
    -o "$OUT/ref_shim_tests" -L"$ROOT/rapidnet_b200/host" -lrapidnet_host -L"$ROOT/rapidnet_b200" -lrapidnet_b200 -lcublas \
    -Xlinker -rpath -Xlinker '$ORIGIN/../../rapidnet_b200/host' -Xlinker -rpath -Xlinker '$ORIGIN/../../rapidnet_b200'
echo "built $OUT/ref_shim_tests"
