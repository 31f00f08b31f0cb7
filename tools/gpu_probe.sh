#!/usr/bin/env bash
# Builds and runs the micro-benchmarks under tools/probes on the GPU box.  Usage (under gpurun): bash tools/gpu_probe.sh <tag>
set -uo pipefail
TAG="${1:-probe}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
cd tools/probes
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tile_gemm_probe tile_gemm_probe.cu && timeout 120 ./tile_gemm_probe 200 | tee "../../$OUT/tile_gemm_probe.txt"
