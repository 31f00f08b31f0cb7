// stream_probe.cu -- micro-benchmark behind the design of the factor-stream phase: how fast can one
// CTA per SM pull a >L2 buffer through a TMA (cp.async.bulk) ring, as a function of ring depth / stage size /
// consumer work?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: consumers only wait/arrive.  mode 1: column-major GEMV work, thread = (row, column group) like k_stream.
// mode 2: GEMV work, warp = column, lane = rows lane+32k (4 accumulators), cross-warp reduce at the end of a unit.
template <int MODE>
__global__ void __launch_bounds__(544, 1) k_ring(const float *__restrict__ src, size_t floats_per_cta, int stages, int stage_floats,
                                                  int nv, float *out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + (size_t)stages * stage_floats);
    uint64_t *empty = full + stages;
    float *wv = reinterpret_cast<float *>(empty + stages);   // 256 floats
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 16); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 256) wv[tid] = 1.0f + tid * 1e-3f;
    __syncthreads();
    const float *base = src + (size_t)blockIdx.x * floats_per_cta;
    const int chunks = (int)(floats_per_cta / stage_floats);
    if (warp == 16) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int c = 0; c < chunks; c++) {
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], stage_floats * 4);
                bulk_g2s(ring + (size_t)st * stage_floats, base + (size_t)c * stage_floats, stage_floats * 4, &full[st]);
                if (++st == stages) { st = 0; ph ^= 1; }
            }
        }
        return;
    }
    int st = 0; uint32_t ph = 0;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int cols = stage_floats / nv;
    const int slots = (nv + 31) & ~31, G = 512 / slots, g = tid / slots, r0 = tid - g * slots;
    for (int c = 0; c < chunks; c++) {
        mbar_wait(&full[st], ph);
        const float *sb = ring + (size_t)st * stage_floats;
        if (MODE == 1) {
            if (g < G && r0 < nv) {
#pragma unroll 4
                for (int j = g; j < cols; j += G) acc0 = fmaf(sb[j * nv + r0], wv[j], acc0);
            }
        } else if (MODE == 2) {
            for (int j = warp; j < cols; j += 16) {
                const float w = wv[j];
                const float *col = sb + j * nv;
                acc0 = fmaf(col[lane], w, acc0);
                acc1 = fmaf(col[lane + 32], w, acc1);
                acc2 = fmaf(col[lane + 64], w, acc2);
                if (lane + 96 < nv) acc3 = fmaf(col[lane + 96], w, acc3);
            }
        } else if (MODE == 3) {   // flat float4 reads, weights ignored: upper bound on smem-read rate
            const float4 *s4 = reinterpret_cast<const float4 *>(sb);
            for (int k = tid; k < stage_floats / 4; k += 512) { const float4 v = s4[k]; acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (MODE != 0) out[blockIdx.x * 512 + tid] = acc0 + acc1 + acc2 + acc3;
}

// plain grid-stride LDG.128 read: what the memory system gives to a simple kernel
__global__ void k_ldg(const float4 *__restrict__ src, size_t n4, float *out) {
    float a = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        a += v0.x + v1.y + v2.z + v3.w;
    }
    if (a == 12345.678f) out[0] = a;
}

template <int MODE>
static double run_ring(const float *d, size_t total_floats, int sms, int stages, int stage_floats, int nv, float *out, int reps) {
    size_t per_cta = total_floats / sms / stage_floats * stage_floats;
    size_t smem = (size_t)stages * stage_floats * 4 + 2 * stages * 8 + 1024 + 128;
    CK(cudaFuncSetAttribute(k_ring<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_ring<MODE><<<sms, 544, smem>>>(d, per_cta, stages, stage_floats, nv, out);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) k_ring<MODE><<<sms, 544, smem>>>(d, per_cta, stages, stage_floats, nv, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return (double)per_cta * sms * 4 * reps / (ms * 1e-3) / 1e9;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const size_t total = (size_t)400 << 20;   // 400 MB > L2
    float *d, *out; CK(cudaMalloc(&d, total)); CK(cudaMalloc(&out, sms * 512 * 4 + 1024));
    CK(cudaMemset(d, 0, total));
    printf("device %s, %d SMs\n", p.name, sms);
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int blocks : {sms * 4, sms * 8, sms * 16}) {
            k_ldg<<<blocks, 256>>>((const float4 *)d, total / 16, out);
            CK(cudaEventRecord(e0));
            for (int r = 0; r < 10; r++) k_ldg<<<blocks, 256>>>((const float4 *)d, total / 16, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("ldg128 read   blocks=%5d            : %8.1f GB/s\n", blocks, (double)total * 10 / (ms * 1e-3) / 1e9);
        }
    }
    const int nv = 97;
    for (int stage_kb : {8, 16, 32}) {
        for (int stages : {3, 4, 6, 8, 12}) {
            const int stage_floats = stage_kb * 1024 / 4;
            if ((size_t)stages * stage_floats * 4 > 200 * 1024) continue;
            double b0 = run_ring<0>(d, total / 4, sms, stages, stage_floats, nv, out, 10);
            double b1 = run_ring<1>(d, total / 4, sms, stages, stage_floats, nv, out, 10);
            double b2 = run_ring<2>(d, total / 4, sms, stages, stage_floats, nv, out, 10);
            double b3 = run_ring<3>(d, total / 4, sms, stages, stage_floats, nv, out, 10);
            printf("ring stage=%2dKB stages=%2d (%3d KB in flight/SM): wait-only %7.1f  gemv(row,grp) %7.1f  gemv(warp=col) %7.1f  float4-sum %7.1f GB/s\n",
                   stage_kb, stages, stage_kb * stages, b0, b1, b2, b3);
        }
    }
    return 0;
}
