// Utilities.cuh (shim) -- the timer of /root/reference/src/Utilities.cuh:28-31, 283-299 (startTicToc / tic / toc /
// stopTicToc: CUDA events on the default stream, toc returns milliseconds) for main()-style drivers.  The reference's device
// kernels declared in that header are internals of its solver and have no counterpart here.
#pragma once
#include "Configuration.h"
namespace rapidnet_shim {
struct TicToc { cudaEvent_t start = nullptr, stop = nullptr; bool ready = false; };
inline TicToc &tictoc() { static TicToc t; return t; }
}
inline void startTicToc() { auto &t = rapidnet_shim::tictoc(); _CUDA(cudaEventCreate(&t.start)); _CUDA(cudaEventCreate(&t.stop)); t.ready = true; }
inline void tic() { auto &t = rapidnet_shim::tictoc(); if (t.ready) _CUDA(cudaEventRecord(t.start, 0)); }
inline real_t toc() {
    auto &t = rapidnet_shim::tictoc();
    if (!t.ready) return -1;
    float ms = 0;
    _CUDA(cudaEventRecord(t.stop, 0)); _CUDA(cudaEventSynchronize(t.stop)); _CUDA(cudaEventElapsedTime(&ms, t.start, t.stop));
    return ms;
}
inline void stopTicToc() { auto &t = rapidnet_shim::tictoc(); if (t.ready) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); t.ready = false; } }
