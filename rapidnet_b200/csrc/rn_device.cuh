// rn_device.cuh -- device-side helpers shared by the CUDA translation units (PTX wrappers for mbarrier and the
// 1-D bulk TMA copy, small reductions).
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace rn {

// ============================================================================================
// PTX helpers (mbarrier + 1-D bulk TMA)
// ============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// raises the transaction count of the current phase without arriving (a later mbar_expect_tx does the arrive)
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, no tensor map): 16-B aligned src/dst, size multiple of 16; SASS: UBLKCP
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the same with an L2 eviction-priority hint (the 64-bit policy encodings of createpolicy: evict_first / evict_last)
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull, kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// global -> L2 bulk prefetch (no shared-memory destination): 16-B aligned address, size multiple of 16; SASS: UBLKPF.L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) {   // projectionBox (Utilities.cu:237-254)
    if (v < lo) return lo; else if (v > hi) return hi; return v;
}

struct Cand { float a; float v; int idx; };
__device__ __forceinline__ void cand_merge(Cand &x, const Cand &y) {   // Isamax: largest |.|, smallest index on ties
    if (y.a > x.a || (y.a == x.a && y.idx < x.idx)) x = y;
}
__device__ __forceinline__ Cand cand_warp(Cand c) {
    for (int o = 16; o > 0; o >>= 1) {
        Cand y;
        y.a = __shfl_xor_sync(0xffffffffu, c.a, o); y.v = __shfl_xor_sync(0xffffffffu, c.v, o); y.idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
        cand_merge(c, y);
    }
    return c;
}


}  // namespace rn
