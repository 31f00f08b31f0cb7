#!/usr/bin/env bash
# Builds the C++ host side (rapidnet_b200/host): librapidnet_host.so = the reference's class surface over the C ABI
# (loaders.cpp, engine.cpp) and host_tests = the reference's test suite replayed against it.  Needs the CUDA library
# (tools/build_lib.sh) and a rapidjson header tree (third-party, MIT; the image ships one inside site-packages).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="$ROOT/rapidnet_b200/host"
CXX="${CXX:-/usr/bin/g++}"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
SITE="$(python -c "import sysconfig; print(sysconfig.get_paths()['purelib'])" 2>/dev/null || true)"
RJ="${RAPIDJSON_INCLUDE:-}"
for d in "$RJ" "$SITE/tilelang/3rdparty/composable_kernel/include" /usr/include /usr/local/include; do
    if [ -n "$d" ] && [ -f "$d/rapidjson/document.h" ]; then RJ="$d"; break; fi
done
if [ -z "$RJ" ] || [ ! -f "$RJ/rapidjson/document.h" ]; then echo "build_host: no rapidjson headers found (set RAPIDJSON_INCLUDE)"; exit 1; fi
FLAGS=(-std=c++17 -O2 -fPIC -Wall -Wno-class-memaccess -I"$ROOT/include" -isystem "$RJ" -I"$CUDA/include")
# hidden inlines + hidden rapidjson (loaders.cpp): the library keeps ITS rapidjson (header-only, inline templates) even when the program
# that loads it was compiled against another rapidjson release (the reference vendors v1.1; interposed inline symbols would crash)
"$CXX" "${FLAGS[@]}" -fvisibility-inlines-hidden -shared -o "$SRC/librapidnet_host.so" "$SRC/loaders.cpp" "$SRC/engine.cpp" \
    -L"$ROOT/rapidnet_b200" -lrapidnet_b200 -L"$CUDA/lib64" -lcublas -lcudart -Wl,-rpath,'$ORIGIN/..'
"$CXX" "${FLAGS[@]}" -o "$SRC/host_tests" "$SRC/host_tests.cpp" -L"$SRC" -lrapidnet_host -L"$ROOT/rapidnet_b200" -lrapidnet_b200 \
    -L"$CUDA/lib64" -lcudart -pthread -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/..'
echo "built $SRC/librapidnet_host.so $SRC/host_tests"
