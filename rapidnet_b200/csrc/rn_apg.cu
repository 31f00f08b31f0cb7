// rn_apg.cu -- the APG iteration of SmpcController on sm_100a.
//
// Reference hot loop: /root/reference/src/SmpcController.cu:1500-1525 (algorithmApg) calling
//   dualExtrapolationStep :535-557, solveStep :563-755, proximalFunG :759-835,
//   computeFixedPointResidual :839-850, dualUpdate :854-864, updatePrimalInfeasibity :1480-1496
// -- about 430 cuBLAS/memcpy/kernel launches, 10 cudaMalloc/cudaFree and 4 host syncs per iteration.
//
// Here one iteration is ("kernels"):
//   k_stream      node-parallel: w = (1+l) y_k - l y_{k-1} fused into the load of the duals, then the
//                 per-node Engine factor matrices are streamed ONCE through a TMA (cp.async.bulk) ring:
//                 a = D xi_w + F psi_w, b = Phi xi_w + Psi psi_w, c = sysF' xi_w.   <- the HBM-roofline kernel
//   k_bwd_chain   one CTA per scenario chain below the last branching stage, Omega/Theta/G in shared memory:
//                 sigma = beta + r_child, r = sigma + a + G q_child, v = -1/2 Omega sigma + Theta q_child + b,
//                 q = c + q_child
//   k_bwd_stage   the same for one branching stage (children summed in-kernel: solveSumChildren)
//   k_fwd_stage / k_fwd_chain   u = ((uhat + u_par) - uhat_par) + L v, x = (x_par + e) + B u, Hx, Hu and the
//                 box projections of the prox, with per-CTA partial sums of the two global distances
//   k_finalize    distance branch (rarely taken), res = Hx - z, y_{k+1} = w + step*res, arg-max-abs log
// captured once in a CUDA graph and replayed `iterations` times; lambda comes from a device table indexed by
// a device-side iteration counter, and y_k / y_{k-1} swap roles by parity instead of being copied.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "rn_internal.h"
#include "rn_device.cuh"

namespace rn {

// ============================================================================================
// k_stream
// ============================================================================================
constexpr int kConsumers = 512;                  // consumer threads (16 warps) + 1 producer warp
constexpr int kStreamThreads = kConsumers + 32;
constexpr int kStages = 6;                       // ring depth
constexpr int kStageFloats = 4096;               // payload per stage (16 KB)

struct StreamArgs {
    const float *mat[4];       // D, F, Phi, Psi (packed per node, Engine.cu:201-207)
    const float *yA_xi, *yA_psi, *yB_xi, *yB_psi;   // dual iterates; parity of the iteration picks y_k / y_{k-1}
    float *w_xi, *w_psi;       // accelerated dual (devVecAcceleratedXi/Psi)
    const float *diag;         // [nodes][ny] s_x | s_xs | s_u
    float *out[2];             // a (D,F products), b (Phi,Psi products)
    float *c;                  // sysF' xi_w
    const float *lambda_tab;
    const int *iter;
    int nodes, nx, nu, nv;
    int halves;                // 2: {D,F} and {Phi,Psi};  1: {D,F} only
    int cols_per_chunk;
    int stage_stride;          // floats per ring stage (payload + 32 floats of slack for the 16-B window)
    int extrapolate;           // 1: w from y_k, y_{k-1};  0: w is read from w_xi / w_psi (step API)
    int dry;                   // profiling: lambda = 0, no vector writes
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

template <int RPT>
__global__ void __launch_bounds__(kStreamThreads, 1) k_stream(const __grid_constant__ StreamArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nx = A.nx, nu = A.nu, nv = A.nv, ny = 2 * nx + nu;
    float *stage_buf = reinterpret_cast<float *>(smem_raw);
    float *wbuf = stage_buf + kStages * A.stage_stride;
    float *red = wbuf + ((ny + 31) & ~31);
    uint64_t *full = reinterpret_cast<uint64_t *>(red + kConsumers);
    uint64_t *empty = full + kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = A.nodes * A.halves;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumers / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kConsumers / 32) {
        // ---------------- producer: one lane drives the TMA ring, running ahead across units ----------------
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int node = u / A.halves, half = u - node * A.halves;
                for (int seg = 0; seg < 2; seg++) {
                    const int ncols = seg == 0 ? 2 * nx : nu;
                    const float *base = A.mat[half * 2 + seg] + (size_t)node * nv * ncols;
                    for (int c0 = 0; c0 < ncols; c0 += A.cols_per_chunk) {
                        const int cc = min(A.cols_per_chunk, ncols - c0);
                        const uintptr_t p0 = reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv);
                        const uintptr_t p1 = p0 + (size_t)cc * nv * sizeof(float);
                        const uintptr_t b0 = p0 & ~uintptr_t(15), b1 = (p1 + 15) & ~uintptr_t(15);
                        const uint32_t bytes = (uint32_t)(b1 - b0);
                        mbar_wait(&empty[st], ph ^ 1);
                        mbar_expect_tx(&full[st], bytes);
                        bulk_g2s(stage_buf + st * A.stage_stride, reinterpret_cast<const void *>(b0), bytes, &full[st]);
                        if (++st == kStages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    int slots = (nv + 31) & ~31;
    if (RPT > 1 || slots > kConsumers) slots = kConsumers;
    const int G = RPT > 1 ? 1 : kConsumers / slots;     // column groups
    const int g = tid / slots, r0 = tid - g * slots;
    const bool active = g < G;
    float lam = 0.f;
    int par = 0;
    if (A.extrapolate && !A.dry) { const int it = *A.iter; lam = A.lambda_tab[it]; par = it & 1; }
    const float *yk_xi = par ? A.yB_xi : A.yA_xi, *yk_psi = par ? A.yB_psi : A.yA_psi;
    const float *ym_xi = par ? A.yA_xi : A.yB_xi, *ym_psi = par ? A.yA_psi : A.yB_psi;
    const float a1 = 1.f + lam, a2 = -lam;

    int st = 0; uint32_t ph = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int node = u / A.halves, half = u - node * A.halves;
        consumer_bar();   // previous unit is done with wbuf / red
        for (int t = tid; t < ny; t += kConsumers) {
            const bool isx = t < 2 * nx;
            const size_t off = isx ? (size_t)node * 2 * nx + t : (size_t)node * nu + (t - 2 * nx);
            float w;
            if (A.extrapolate) {
                const float y = isx ? yk_xi[off] : yk_psi[off];
                const float yp = isx ? ym_xi[off] : ym_psi[off];
                w = y * a1;            // Sscal   (SmpcController.cu:548-549)
                w += a2 * yp;          // Saxpy   (:551-552)
                if (half == 0 && !A.dry) { if (isx) A.w_xi[off] = w; else A.w_psi[off] = w; }
            } else {
                w = isx ? A.w_xi[off] : A.w_psi[off];
            }
            wbuf[t] = w;
        }
        consumer_bar();
        if (half == 0 && !A.dry) {   // c = sysF' xi_w = s_x . xi_1 + s_xs . xi_2   (:651-658)
            const float *dg = A.diag + (size_t)node * ny;
            for (int t = tid; t < nx; t += kConsumers) A.c[(size_t)node * nx + t] = dg[t] * wbuf[t] + dg[nx + t] * wbuf[nx + t];
        }
        float acc[RPT];
#pragma unroll
        for (int k = 0; k < RPT; k++) acc[k] = 0.f;
        for (int seg = 0; seg < 2; seg++) {
            const int ncols = seg == 0 ? 2 * nx : nu;
            const float *base = A.mat[half * 2 + seg] + (size_t)node * nv * ncols;
            const float *wseg = wbuf + (seg == 0 ? 0 : 2 * nx);
            for (int c0 = 0; c0 < ncols; c0 += A.cols_per_chunk) {
                const int cc = min(A.cols_per_chunk, ncols - c0);
                const int off = (int)((reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv) & 15) >> 2);
                mbar_wait(&full[st], ph);
                const float *sb = stage_buf + st * A.stage_stride + off;
                if (active) {
#pragma unroll 4
                    for (int j = g; j < cc; j += G) {
                        const float wv = wseg[c0 + j];
                        const float *col = sb + j * nv;
#pragma unroll
                        for (int k = 0; k < RPT; k++) {
                            const int r = r0 + k * kConsumers;
                            if (r < nv) acc[k] = fmaf(col[r], wv, acc[k]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                if (++st == kStages) { st = 0; ph ^= 1; }
            }
        }
        float *out = A.out[half];
        if (G > 1) {
            if (active && r0 < nv) red[g * nv + r0] = acc[0];
            consumer_bar();
            if (tid < nv && !A.dry) {
                float s = red[tid];
                for (int gg = 1; gg < G; gg++) s += red[gg * nv + tid];
                out[(size_t)node * nv + tid] = s;
            }
        } else if (!A.dry) {
#pragma unroll
            for (int k = 0; k < RPT; k++) {
                const int r = r0 + k * kConsumers;
                if (r < nv) out[(size_t)node * nv + r] = acc[k];
            }
        }
    }
}

// ring stage payload: 16 KB; 32 KB for wide matrices (nv > 128: a 16 KB stage would hold ten columns of the 4 x network)
static int stream_stage_floats(const Handle *h) { return h->d.nv > 128 ? 2 * kStageFloats : kStageFloats; }
static size_t stream_smem_bytes(const Handle *h) {
    const int ny = 2 * h->d.nx + h->d.nu;
    return (size_t)kStages * (stream_stage_floats(h) + 32) * 4 + (size_t)((ny + 31) & ~31) * 4 + kConsumers * 4 + 2 * kStages * 8 + 64;
}

// ============================================================================================
// block-level helpers for the sweeps
// ============================================================================================
// ys[r] = sum_c A[r + c*lda] * xs[c], r < m.  A in global or shared memory, xs/ys/scratch in shared memory.
// scratch needs NT floats.  Ends with a __syncthreads().
template <int NT>
__device__ __forceinline__ void block_gemv(const float *__restrict__ A, int lda, int m, int n, const float *xs, float *ys,
                                           float *scratch) {
    const int t = threadIdx.x;
    int slots = (m + 31) & ~31;
    if (slots > NT) slots = NT;
    const int G = NT / slots;
    if (G <= 1) {
        for (int r = t; r < m; r += NT) {
            float acc = 0.f;
#pragma unroll 4
            for (int c = 0; c < n; c++) acc = fmaf(A[r + (size_t)c * lda], xs[c], acc);
            ys[r] = acc;
        }
        __syncthreads();
    } else {
        const int g = t / slots, rr = t - g * slots;
        if (g < G && rr < m) {
            float acc = 0.f;
#pragma unroll 4
            for (int c = g; c < n; c += G) acc = fmaf(A[rr + (size_t)c * lda], xs[c], acc);
            scratch[g * m + rr] = acc;
        }
        __syncthreads();
        if (t < m) {
            float s = scratch[t];
            for (int gg = 1; gg < G; gg++) s += scratch[gg * m + t];
            ys[t] = s;
        }
        __syncthreads();
    }
}

struct SweepArgs {
    // tree
    const int *parent, *child_first, *child_count, *omega_idx;
    // constants
    const float *Omega, *Theta, *G, *L, *B, *diag;
    const float *beta, *uhat, *e, *xcur, *uprev, *uhat_prev;
    const float *sxmin, *sxmax, *sxs, *sumin, *sumax;
    // hoisted products and state
    const float *a, *b, *c;
    float *q, *r, *sigma, *V, *U, *X;
    const float *w_xi, *w_psi;
    float *pri_xi, *pri_psi, *dual_xi, *dual_psi;
    double *dist_part;
    int nx, nu, nv;
    int df_mode;       // v = -1/2 Omega r
    int fuse_prox;     // forward kernels also do t = Hx + w/step, z = clamp(t), distance partials
    float inv_step;
};

// Hx, Hu (SmpcController.cu:744-747) and, if fused, the box part of proximalFunG (:778-789, :827) for one node.
// xs/us: x_i and u_i in shared memory.  Returns this thread's partial sums of squares.
template <int NT>
__device__ __forceinline__ void node_epilogue(const SweepArgs &S, int i, const float *xs, const float *us, double &s1,
                                              double &s2) {
    const int nx = S.nx, nu = S.nu, ny = 2 * nx + nu;
    const float *dg = S.diag + (size_t)i * ny;
    for (int t = threadIdx.x; t < ny; t += NT) {
        if (t < 2 * nx) {
            const int j = t < nx ? t : t - nx;
            const size_t k = (size_t)i * 2 * nx + t;
            const float hx = dg[t] * xs[j];
            S.pri_xi[k] = hx;
            if (S.fuse_prox) {
                const float tt = hx + S.inv_step * S.w_xi[k];
                const size_t kb = (size_t)i * nx + j;
                const float z = t < nx ? clampf(tt, S.sxmin[kb], S.sxmax[kb]) : clampf(tt, S.sxs[kb], __int_as_float(0x7F7F7F7F));
                S.dual_xi[k] = z;
                const float df = tt + -1.f * z;
                if (t < nx) s1 += (double)df * df; else s2 += (double)df * df;
            }
        } else {
            const int j = t - 2 * nx;
            const size_t k = (size_t)i * nu + j;
            const float hu = dg[t] * us[j];
            S.pri_psi[k] = hu;
            if (S.fuse_prox) S.dual_psi[k] = clampf(hu + S.inv_step * S.w_psi[k], S.sumin[k], S.sumax[k]);
        }
    }
}

template <int NT>
__device__ __forceinline__ void block_store_dist(double s1, double s2, double *dst, double *sh /* 2*NT/32 doubles */) {
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[warp] = s1; sh[NT / 32 + warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0, t2 = 0;
        for (int w = 0; w < NT / 32; w++) { t1 += sh[w]; t2 += sh[NT / 32 + w]; }
        dst[0] = t1; dst[1] = t2;
    }
}

// ============================================================================================
// per-stage sweeps (branching stages; every stage in RN_SWEEP_PER_STAGE)
// ============================================================================================
constexpr int kStageThreads = 256;

// one CTA per node of the stage: solveSumChildren (Utilities.cu:168-201) + the eight batched GEMVs of
// SmpcController.cu:597-659, with the per-node products a, b, c already formed by k_stream.
__global__ void __launch_bounds__(kStageThreads) k_bwd_stage(const __grid_constant__ SweepArgs S, int first) {
    extern __shared__ __align__(16) float sm[];
    const int nx = S.nx, nv = S.nv, i = first + blockIdx.x, t = threadIdx.x;
    float *qs = sm, *sg = qs + nx, *y1 = sg + nv, *y2 = y1 + nv, *y3 = y2 + nv, *scr = y3 + nv;
    const int c0 = S.child_first[i], nc = S.child_count[i];
    for (int k = t; k < nx; k += kStageThreads) {
        float s = 0.f;
        if (nc > 0) { s = S.q[(size_t)c0 * nx + k]; for (int c = 1; c < nc; c++) s += S.q[(size_t)(c0 + c) * nx + k]; }
        qs[k] = s;
    }
    for (int k = t; k < nv; k += kStageThreads) {
        float s = S.beta[(size_t)i * nv + k];
        if (nc > 0) { float rs = S.r[(size_t)c0 * nv + k]; for (int c = 1; c < nc; c++) rs += S.r[(size_t)(c0 + c) * nv + k]; s += rs; }
        sg[k] = s;
        S.sigma[(size_t)i * nv + k] = s;
    }
    __syncthreads();
    const int oi = S.omega_idx[i];
    const float *Om = S.Omega + (size_t)oi * nv * nv, *Th = S.Theta + (size_t)oi * nv * nx;
    block_gemv<kStageThreads>(S.G, nv, nv, nx, qs, y3, scr);          // G q      (:644-646)
    for (int k = t; k < nv; k += kStageThreads) y3[k] = sg[k] + S.a[(size_t)i * nv + k] + y3[k];   // r = sigma + D xi + F psi + G q
    __syncthreads();
    if (S.df_mode) {
        block_gemv<kStageThreads>(Om, nv, nv, nv, y3, y1, scr);       // v = -1/2 Omega r
        for (int k = t; k < nv; k += kStageThreads) { S.V[(size_t)i * nv + k] = -0.5f * y1[k]; S.r[(size_t)i * nv + k] = y3[k]; }
    } else {
        block_gemv<kStageThreads>(Om, nv, nv, nv, sg, y1, scr);       // Omega sigma (:604-607)
        block_gemv<kStageThreads>(Th, nv, nv, nx, qs, y2, scr);       // Theta q     (:611-613)
        for (int k = t; k < nv; k += kStageThreads) {
            S.V[(size_t)i * nv + k] = (-0.5f * y1[k] + y2[k]) + S.b[(size_t)i * nv + k];
            S.r[(size_t)i * nv + k] = y3[k];
        }
    }
    for (int k = t; k < nx; k += kStageThreads) S.q[(size_t)i * nx + k] = S.c[(size_t)i * nx + k] + qs[k];   // q = F' xi + q
}

// one CTA per node of the stage: forward substitution (SmpcController.cu:678-741) + Hx (:744-747) [+ prox boxes]
__global__ void __launch_bounds__(kStageThreads) k_fwd_stage(const __grid_constant__ SweepArgs S, int first, int branching,
                                                             int slot0) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double dsh[2 * kStageThreads / 32];
    const int nx = S.nx, nu = S.nu, nv = S.nv, i = first + blockIdx.x, t = threadIdx.x;
    float *vs = sm, *us = vs + nv, *lv = us + nu, *xs = lv + nu, *bu = xs + nx, *scr = bu + nx;
    const int par = S.parent[i];
    for (int k = t; k < nv; k += kStageThreads) vs[k] = S.V[(size_t)i * nv + k];
    __syncthreads();
    block_gemv<kStageThreads>(S.L, nu, nu, nv, vs, lv, scr);          // L v
    for (int k = t; k < nu; k += kStageThreads) {
        const float uh = S.uhat[(size_t)i * nu + k];
        const float up = par < 0 ? S.uprev[k] : S.U[(size_t)par * nu + k];
        const float uhp = par < 0 ? S.uhat_prev[k] : S.uhat[(size_t)par * nu + k];
        float u;
        if (branching) u = (up + -1.f * uhp) + (uh + lv[k]);          // :701-710
        else u = ((uh + up) + -1.f * uhp) + lv[k];                    // :683-693, :722-728
        us[k] = u;
        S.U[(size_t)i * nu + k] = u;
    }
    __syncthreads();
    block_gemv<kStageThreads>(S.B, nx, nx, nu, us, bu, scr);          // B u
    for (int k = t; k < nx; k += kStageThreads) {
        const float xp = par < 0 ? S.xcur[k] : S.X[(size_t)par * nx + k];
        const float ei = S.e[(size_t)i * nx + k];
        const float x = branching ? xp + (ei + bu[k]) : (xp + ei) + bu[k];   // :712-719 / :730-737
        xs[k] = x;
        S.X[(size_t)i * nx + k] = x;
    }
    __syncthreads();
    double s1 = 0, s2 = 0;
    node_epilogue<kStageThreads>(S, i, xs, us, s1, s2);
    if (S.fuse_prox) block_store_dist<kStageThreads>(s1, s2, S.dist_part + 2 * (size_t)(slot0 + blockIdx.x), dsh);
}

static size_t bwd_stage_smem(const Handle *h) { return (size_t)(h->d.nx + 4 * h->d.nv + kStageThreads) * 4; }
static size_t fwd_stage_smem(const Handle *h) { return (size_t)(h->d.nv + 2 * h->d.nu + 2 * h->d.nx + kStageThreads) * 4; }

// ============================================================================================
// chain sweeps: one CTA per scenario below the last branching stage, shared matrices in smem
// ============================================================================================
constexpr int kChainThreads = 512;

__global__ void __launch_bounds__(kChainThreads) k_bwd_chain(const __grid_constant__ SweepArgs S, const int *__restrict__ cum,
                                                             int stage_first, int stage_last) {
    extern __shared__ __align__(16) float sm[];
    const int nx = S.nx, nv = S.nv, j = blockIdx.x, t = threadIdx.x;
    float *Om = sm, *Th = Om + nv * nv, *Gm = Th + nv * nx;
    float *qs = Gm + nv * nx, *sg = qs + nx, *rv = sg + nv, *y1 = rv + nv, *y2 = y1 + nv, *y3 = y2 + nv, *scr = y3 + nv;
    {
        const int oi = S.omega_idx[cum[stage_last] + j];   // the whole chain aliases one Omega/Theta (Engine.cu:210-221)
        const float *gO = S.Omega + (size_t)oi * nv * nv, *gT = S.Theta + (size_t)oi * nv * nx;
        for (int k = t; k < nv * nv; k += kChainThreads) Om[k] = gO[k];
        for (int k = t; k < nv * nx; k += kChainThreads) { Th[k] = gT[k]; Gm[k] = S.G[k]; }
        for (int k = t; k < nx; k += kChainThreads) qs[k] = 0.f;
        for (int k = t; k < nv; k += kChainThreads) rv[k] = 0.f;
    }
    __syncthreads();
    for (int s = stage_last; s >= stage_first; s--) {
        const int i = cum[s] + j;
        // operands of this stage, fetched before the shared-memory GEMVs need them
        float be = 0.f, av = 0.f, bv = 0.f, cv = 0.f;
        if (t < nv) { be = S.beta[(size_t)i * nv + t]; av = S.a[(size_t)i * nv + t]; if (!S.df_mode) bv = S.b[(size_t)i * nv + t]; }
        if (t < nx) cv = S.c[(size_t)i * nx + t];
        if (t < nv) { const float sgv = be + rv[t]; sg[t] = sgv; S.sigma[(size_t)i * nv + t] = sgv; }
        __syncthreads();
        block_gemv<kChainThreads>(Gm, nv, nv, nx, qs, y3, scr);
        if (t < nv) { const float rr = sg[t] + av + y3[t]; rv[t] = rr; S.r[(size_t)i * nv + t] = rr; }
        __syncthreads();
        if (S.df_mode) {
            block_gemv<kChainThreads>(Om, nv, nv, nv, rv, y1, scr);
            if (t < nv) S.V[(size_t)i * nv + t] = -0.5f * y1[t];
        } else {
            block_gemv<kChainThreads>(Om, nv, nv, nv, sg, y1, scr);
            block_gemv<kChainThreads>(Th, nv, nv, nx, qs, y2, scr);
            if (t < nv) S.V[(size_t)i * nv + t] = (-0.5f * y1[t] + y2[t]) + bv;
        }
        if (t < nx) { const float qq = cv + qs[t]; qs[t] = qq; S.q[(size_t)i * nx + t] = qq; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kChainThreads) k_fwd_chain(const __grid_constant__ SweepArgs S, const int *__restrict__ cum,
                                                             int stage_first, int stage_last, int slot0) {
    extern __shared__ __align__(16) float sm[];
    __shared__ double dsh[2 * kChainThreads / 32];
    const int nx = S.nx, nu = S.nu, nv = S.nv, j = blockIdx.x, t = threadIdx.x;
    float *Ls = sm, *Bs = Ls + nu * nv, *vs = Bs + nx * nu, *us = vs + nv, *lv = us + nu, *xs = lv + nu, *bu = xs + nx, *scr = bu + nx;
    for (int k = t; k < nu * nv; k += kChainThreads) Ls[k] = S.L[k];
    for (int k = t; k < nx * nu; k += kChainThreads) Bs[k] = S.B[k];
    const int par0 = S.parent[cum[stage_first] + j];
    float up = 0.f, uhp = 0.f;          // u and uhat of the parent, element t (t < nu)
    if (t < nu) {
        up = par0 < 0 ? S.uprev[t] : S.U[(size_t)par0 * nu + t];
        uhp = par0 < 0 ? S.uhat_prev[t] : S.uhat[(size_t)par0 * nu + t];
    }
    if (t < nx) xs[t] = par0 < 0 ? S.xcur[t] : S.X[(size_t)par0 * nx + t];
    double s1 = 0, s2 = 0;
    __syncthreads();
    for (int s = stage_first; s <= stage_last; s++) {
        const int i = cum[s] + j;
        float uh = 0.f, ei = 0.f;
        if (t < nu) uh = S.uhat[(size_t)i * nu + t];
        if (t < nx) ei = S.e[(size_t)i * nx + t];
        if (t < nv) vs[t] = S.V[(size_t)i * nv + t];
        __syncthreads();
        block_gemv<kChainThreads>(Ls, nu, nu, nv, vs, lv, scr);
        if (t < nu) {
            const float u = ((uh + up) + -1.f * uhp) + lv[t];      // a chain stage never branches (:722-728)
            us[t] = u; S.U[(size_t)i * nu + t] = u;
            up = u; uhp = uh;
        }
        __syncthreads();
        block_gemv<kChainThreads>(Bs, nx, nx, nu, us, bu, scr);
        if (t < nx) { const float x = (xs[t] + ei) + bu[t]; xs[t] = x; S.X[(size_t)i * nx + t] = x; }
        __syncthreads();
        node_epilogue<kChainThreads>(S, i, xs, us, s1, s2);
    }
    if (S.fuse_prox) block_store_dist<kChainThreads>(s1, s2, S.dist_part + 2 * (size_t)(slot0 + j), dsh);
}

static size_t bwd_chain_smem(const Handle *h) {
    const size_t nx = h->d.nx, nv = h->d.nv;
    return (nv * nv + 2 * nv * nx + nx + 5 * nv + kChainThreads) * 4;
}
static size_t fwd_chain_smem(const Handle *h) {
    const size_t nx = h->d.nx, nu = h->d.nu, nv = h->d.nv;
    return (nu * nv + nx * nu + nv + 2 * nu + 2 * nx + kChainThreads) * 4;
}

// ============================================================================================
// element-wise kernels
// ============================================================================================
constexpr int kEwThreads = 256;

// dualExtrapolationStep as a stand-alone step (step API only; the fused path does it inside k_stream)
__global__ void k_extrapolate(size_t n_xi, size_t n_psi, float lam, const float *__restrict__ y_xi, const float *__restrict__ y_psi,
                              float *__restrict__ ym_xi, float *__restrict__ ym_psi, float *__restrict__ w_xi, float *__restrict__ w_psi) {
    const float a1 = 1.f + lam, a2 = -lam;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_xi + n_psi; k += (size_t)gridDim.x * blockDim.x) {
        const bool isx = k < n_xi;
        const size_t o = isx ? k : k - n_xi;
        const float y = isx ? y_xi[o] : y_psi[o];
        const float yp = isx ? ym_xi[o] : ym_psi[o];
        float w = y * a1; w += a2 * yp;
        if (isx) { w_xi[o] = w; ym_xi[o] = y; } else { w_psi[o] = w; ym_psi[o] = y; }   // y_{k-1} <- y_k (:554-555)
    }
}

// box part of proximalFunG for all nodes (step API and RN_SWEEP_PER_STAGE when the forward kernels do not fuse it)
__global__ void __launch_bounds__(kEwThreads) k_prox_boxes(const __grid_constant__ SweepArgs S, int nodes) {
    __shared__ double dsh[2 * kEwThreads / 32];
    const int nx = S.nx, nu = S.nu, ny = 2 * nx + nu, i = blockIdx.x;
    if (i >= nodes) return;
    double s1 = 0, s2 = 0;
    for (int t = threadIdx.x; t < ny; t += kEwThreads) {
        if (t < 2 * nx) {
            const int j = t < nx ? t : t - nx;
            const size_t k = (size_t)i * 2 * nx + t, kb = (size_t)i * nx + j;
            const float tt = S.pri_xi[k] + S.inv_step * S.w_xi[k];
            const float z = t < nx ? clampf(tt, S.sxmin[kb], S.sxmax[kb]) : clampf(tt, S.sxs[kb], __int_as_float(0x7F7F7F7F));
            S.dual_xi[k] = z;
            const float df = tt + -1.f * z;
            if (t < nx) s1 += (double)df * df; else s2 += (double)df * df;
        } else {
            const size_t k = (size_t)i * nu + (t - 2 * nx);
            S.dual_psi[k] = clampf(S.pri_psi[k] + S.inv_step * S.w_psi[k], S.sumin[k], S.sumax[k]);
        }
    }
    block_store_dist<kEwThreads>(s1, s2, S.dist_part + 2 * (size_t)i, dsh);
}

struct FinalArgs {
    const float *pri_xi, *pri_psi, *w_xi, *w_psi;
    float *dual_xi, *dual_psi, *res_xi, *res_psi;
    float *yA_xi, *yA_psi, *yB_xi, *yB_psi;   // fused: y_{k+1} goes where y_{k-1} lived (parity);  step API: yA = update buffers
    const double *dist_part;
    float *scal;          // [0] d1, [1] d2, [2] branch-1 taken, [3] branch-2 taken
    float *pinf;          // per-iteration primal infeasibility log
    float *pinf_part;     // [grid][6] : abs, signed, idx for xi ; abs, signed, idx for psi
    float *pinf4;         // nullable: [iterations][4] |res|, res at the arg-max of the xi block and of the psi block (rn_read_pinf_parts)
    int *iter;
    unsigned int *done;
    int nodes, nx, nu, n_slots;
    int do_branch, do_residual, do_update, parity_swap, log_inf;
    float step, inv_step, pen_x, pen_xs;
};

// distance branch of proximalFunG (:792-797, :810-815; quirk SURVEY A.4-1), computeFixedPointResidual (:839-850),
// dualUpdate (:854-864) and updatePrimalInfeasibity (:1480-1496) in one pass over the duals.
__global__ void __launch_bounds__(kEwThreads) k_finalize(const __grid_constant__ FinalArgs F) {
    __shared__ double dsh[2 * kEwThreads / 32];
    __shared__ float sd[2];
    __shared__ Cand csh[2 * kEwThreads / 32];
    __shared__ bool is_last;
    const int nx = F.nx, nu = F.nu, ny = 2 * nx + nu, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    float d1 = 0.f, d2 = 0.f;
    if (F.do_branch) {
        // every CTA reduces the same partials in the same order -> identical, deterministic distances
        double s1 = 0, s2 = 0;
        for (int k = t; k < F.n_slots; k += kEwThreads) { s1 += F.dist_part[2 * (size_t)k]; s2 += F.dist_part[2 * (size_t)k + 1]; }
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) { dsh[warp] = s1; dsh[kEwThreads / 32 + warp] = s2; }
        __syncthreads();
        if (t == 0) {
            double t1 = 0, t2 = 0;
            for (int w = 0; w < kEwThreads / 32; w++) { t1 += dsh[w]; t2 += dsh[kEwThreads / 32 + w]; }
            sd[0] = (float)sqrt(t1); sd[1] = (float)sqrt(t2);   // cublasSnrm2 (:792, :810)
        }
        __syncthreads();
        d1 = sd[0]; d2 = sd[1];
    }
    const float thr1 = F.inv_step * F.pen_x, thr2 = F.inv_step * F.pen_xs;
    const bool br1 = F.do_branch && d1 > thr1, br2 = F.do_branch && d2 > thr2;
    const float sc1 = br1 ? 1.f - thr1 / d1 : 0.f, sc2 = br2 ? 1.f - thr2 / d2 : 0.f;
    int par = 0;
    if (F.parity_swap) par = (*F.iter) & 1;
    float *yn_xi = par ? F.yA_xi : F.yB_xi, *yn_psi = par ? F.yA_psi : F.yB_psi;   // y_{k+1} overwrites y_{k-1}
    if (!F.parity_swap) { yn_xi = F.yA_xi; yn_psi = F.yA_psi; }
    Cand bx{-1.f, 0.f, 0x7fffffff}, bp{-1.f, 0.f, 0x7fffffff};
    for (int i = blockIdx.x; i < F.nodes; i += gridDim.x) {
        for (int e = t; e < ny; e += kEwThreads) {
            const bool isx = e < 2 * nx;
            const size_t k = isx ? (size_t)i * 2 * nx + e : (size_t)i * nu + (e - 2 * nx);
            const float hx = isx ? F.pri_xi[k] : F.pri_psi[k];
            const float w = isx ? F.w_xi[k] : F.w_psi[k];
            float z = isx ? F.dual_xi[k] : F.dual_psi[k];
            if (isx && (br1 || br2)) {
                const float tt = hx + F.inv_step * w;
                const float df = tt + -1.f * z;
                if (e < nx) { if (br1) { z = z + sc1 * df; F.dual_xi[k] = z; } }
                else if (br2) {
                    // when branch 1 fired the reference has already clobbered its scratch (:800-802): node 0 sees 0,
                    // every other node sees (t - z) - z in the safety half
                    const float d2v = br1 ? (i == 0 ? 0.f : df + -1.f * z) : df;
                    z = z + sc2 * d2v; F.dual_xi[k] = z;
                }
            }
            float res;
            if (F.do_residual) { res = hx + -1.f * z; if (isx) F.res_xi[k] = res; else F.res_psi[k] = res; }
            else res = isx ? F.res_xi[k] : F.res_psi[k];
            if (F.do_update) { const float yn = w + F.step * res; if (isx) yn_xi[k] = yn; else yn_psi[k] = yn; }
            if (F.log_inf) {
                Cand c{fabsf(res), res, (int)k};
                if (isx) cand_merge(bx, c); else cand_merge(bp, c);
            }
        }
    }
    if (!F.log_inf) {
        if (F.do_branch && blockIdx.x == 0 && t == 0) { F.scal[0] = d1; F.scal[1] = d2; F.scal[2] = br1; F.scal[3] = br2; }
        return;
    }
    bx = cand_warp(bx); bp = cand_warp(bp);
    if (lane == 0) { csh[warp] = bx; csh[kEwThreads / 32 + warp] = bp; }
    __syncthreads();
    if (t == 0) {
        Cand x = csh[0], p = csh[kEwThreads / 32];
        for (int w = 1; w < kEwThreads / 32; w++) { cand_merge(x, csh[w]); cand_merge(p, csh[kEwThreads / 32 + w]); }
        float *o = F.pinf_part + 6 * (size_t)blockIdx.x;
        o[0] = x.a; o[1] = x.v; o[2] = __int_as_float(x.idx); o[3] = p.a; o[4] = p.v; o[5] = __int_as_float(p.idx);
        __threadfence();
        const unsigned int prev = atomicAdd(F.done, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && t == 0) {
        __threadfence();
        Cand x{-1.f, 0.f, 0x7fffffff}, p{-1.f, 0.f, 0x7fffffff};
        for (unsigned int b = 0; b < gridDim.x; b++) {
            const volatile float *o = F.pinf_part + 6 * (size_t)b;
            Cand cx{o[0], o[1], __float_as_int(o[2])}, cp{o[3], o[4], __float_as_int(o[5])};
            cand_merge(x, cx); cand_merge(p, cp);
        }
        const int it = *F.iter;
        F.pinf[it] = fmaxf(x.v, p.v);    // max(maxValueXi, maxValuePsi) (:1495)
        if (F.pinf4) { float *o4 = F.pinf4 + 4 * (size_t)it; o4[0] = x.a; o4[1] = x.v; o4[2] = p.a; o4[3] = p.v; }
        F.scal[0] = d1; F.scal[1] = d2; F.scal[2] = br1; F.scal[3] = br2;
        *F.iter = it + 1;
        *F.done = 0u;
    }
}

__global__ void k_clamp_vec(int n, float *__restrict__ v, const float *__restrict__ lo, const float *__restrict__ hi) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) v[t] = clampf(v[t], lo[t], hi[t]);
}

// moveForewardInTime (:1692-1698): stateUpdate = x_cur + B u0 ; the disturbance lands in devVecX
__global__ void k_move_forward(int nx, int nu, const float *__restrict__ xcur, const float *__restrict__ B,
                               const float *__restrict__ u, const float *__restrict__ e, float *__restrict__ X,
                               float *__restrict__ out) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nx; r += gridDim.x * blockDim.x) {
        float acc = xcur[r];
        X[r] += e[r];
        for (int c = 0; c < nu; c++) acc += B[r + (size_t)c * nx] * u[c];
        out[r] = acc;
    }
}

// ============================================================================================
// host side
// ============================================================================================
void fill_lambda_table(std::vector<float> &tab, int iters) {
    // SmpcController.cu:1505-1520: theta is float, the update is evaluated in double (pow/sqrt) and rounded back
    tab.resize(iters);
    float theta0 = 1, theta1 = 1;
    for (int k = 0; k < iters; k++) {
        tab[k] = theta1 * (1 / theta0 - 1);
        theta0 = theta1;
        theta1 = 0.5 * (sqrt(pow(theta1, 4) + 4 * pow(theta1, 2)) - pow(theta1, 2));
    }
}

double stream_bytes_per_iteration(const Handle *h) {
    const rn_dims &d = h->d;
    const double ny = 2.0 * d.nx + d.nu;
    if (h->factor_mode == RN_FACTORS_SHARED) {
        // no per-node matrix: read Hx, w, z, y_{k-1} + write y_k, w + read the diagonals (ny) + write c, G c, L' g; the two
        // shared matrices G, L' once per CTA from L2 (counted once)
        const double vec = 6.0 * ny + ny + d.nx + 2.0 * d.nv;
        return 4.0 * (d.nodes * vec + (double)d.nv * (d.nx + d.nu));
    }
    const double per_node_mats = (double)d.nv * (2.0 * d.nx + d.nu) * (h->factor_mode == RN_FACTORS_FULL ? 2.0 : 1.0);
    // matrices once + read y_k, y_{k-1} + write w + read diag(2nx) + write a,(b),c
    const double vec = 3.0 * ny + 2.0 * d.nx + d.nx + d.nv * (h->factor_mode == RN_FACTORS_FULL ? 2.0 : 1.0);
    return 4.0 * d.nodes * (per_node_mats + vec);
}

double apg_bytes_per_iteration(const Handle *h) {
    // SURVEY 8(d) "Tier A": bytes_iter = 4 [ nodes (nv(4nx+2nu) + (2nx+nu) + V) + fb (nv^2 + nv nx) + nv nx + nu nv + nx nu ]
    // "Tier B" (RN_FACTORS_SHARED): the per-node matrix term vanishes (vectors + diagonals + the shared matrices)
    const rn_dims &d = h->d;
    const double nx = d.nx, nu = d.nu, nv = d.nv, ny = 2 * nx + nu;
    const double V = 4 * ny + ny + 2 * nv + 2 * (nx + nv) + 2 * nu + 2 * nx + ny + 2 * ny + 3 * nx + 2 * nu + 2 * ny;
    const double mats = h->factor_mode == RN_FACTORS_SHARED ? 0.0 : nv * (4 * nx + 2 * nu) * (h->factor_mode == RN_FACTORS_FULL ? 1.0 : 0.5);
    return 4.0 * (d.nodes * (mats + ny + V) + (double)h->n_omega * (nv * nv + nv * nx) + nv * nx + nu * nv + nx * nu);
}

bool use_persistent(const Handle *h);
static rn_status enqueue_persistent(Handle *h, int iterations);
static bool chain_fits(const Handle *h) {
    return h->chain_stage < h->d.N && bwd_chain_smem(h) <= 200 * 1024 && fwd_chain_smem(h) <= 200 * 1024 &&
           h->d.nv <= kChainThreads && h->d.nu <= kChainThreads && h->d.nx <= kChainThreads;
}

static SweepArgs make_sweep_args(Handle *h, bool fuse_prox) {
    SweepArgs S{};
    S.parent = h->t.parent; S.child_first = h->t.child_first; S.child_count = h->t.child_count; S.omega_idx = h->t.omega_idx;
    S.Omega = h->Omega; S.Theta = h->Theta; S.G = h->G; S.L = h->L; S.B = h->B; S.diag = h->diag;
    S.beta = h->beta; S.uhat = h->uhat; S.e = h->e; S.xcur = h->xcur; S.uprev = h->uprev; S.uhat_prev = h->uhat_prev;
    S.sxmin = h->sxmin; S.sxmax = h->sxmax; S.sxs = h->sxs; S.sumin = h->sumin; S.sumax = h->sumax;
    S.a = h->a; S.b = h->b; S.c = h->c; S.q = h->q; S.r = h->r; S.sigma = h->sigma; S.V = h->V; S.U = h->U; S.X = h->X;
    S.w_xi = h->acc_xi; S.w_psi = h->acc_psi; S.pri_xi = h->pri_xi; S.pri_psi = h->pri_psi;
    S.dual_xi = h->dual_xi; S.dual_psi = h->dual_psi; S.dist_part = h->dist_part;
    S.nx = h->d.nx; S.nu = h->d.nu; S.nv = h->d.nv;
    S.df_mode = h->factor_mode == RN_FACTORS_DF ? 1 : 0;
    S.fuse_prox = fuse_prox ? 1 : 0;
    S.inv_step = 1 / h->step;
    return S;
}

static rn_status launch_stream(Handle *h, cudaStream_t st, bool extrapolate, bool dry, const float *yA_xi, const float *yA_psi,
                               const float *yB_xi, const float *yB_psi) {
    const rn_dims &d = h->d;
    if (h->factor_mode == RN_FACTORS_SHARED)
        return fail(h, RN_ERR_INVALID, "the stand-alone factor stream does not exist in RN_FACTORS_SHARED mode (persistent sweep only)");
    StreamArgs A{};
    A.mat[0] = h->D; A.mat[1] = h->F; A.mat[2] = h->Phi; A.mat[3] = h->Psi;
    A.yA_xi = yA_xi; A.yA_psi = yA_psi; A.yB_xi = yB_xi; A.yB_psi = yB_psi;
    A.w_xi = h->acc_xi; A.w_psi = h->acc_psi; A.diag = h->diag;
    A.out[0] = h->a; A.out[1] = h->b; A.c = h->c;
    A.lambda_tab = h->lambda_tab; A.iter = h->iter_dev;
    A.nodes = d.nodes; A.nx = d.nx; A.nu = d.nu; A.nv = d.nv;
    A.halves = h->factor_mode == RN_FACTORS_FULL ? 2 : 1;
    A.cols_per_chunk = stream_stage_floats(h) / d.nv;
    A.stage_stride = stream_stage_floats(h) + 32;
    if (A.cols_per_chunk < 1) return fail(h, RN_ERR_INVALID, "nv = %d exceeds the stream kernel's stage (%d floats)", d.nv, stream_stage_floats(h));
    A.extrapolate = extrapolate ? 1 : 0;
    A.dry = dry ? 1 : 0;
    const size_t smem = stream_smem_bytes(h);
    const int units = A.nodes * A.halves;
    const int grid = std::min(units, h->sm_count);
    const int rpt = ceil_div(d.nv, kConsumers);
    if (rpt == 1) {
        RN_CUDA(h, cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_stream<1><<<grid, kStreamThreads, smem, st>>>(A);
    } else if (rpt == 2) {
        RN_CUDA(h, cudaFuncSetAttribute(k_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_stream<2><<<grid, kStreamThreads, smem, st>>>(A);
    } else if (rpt <= 4) {
        RN_CUDA(h, cudaFuncSetAttribute(k_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_stream<4><<<grid, kStreamThreads, smem, st>>>(A);
    } else {
        return fail(h, RN_ERR_INVALID, "nv = %d not supported by the stream kernel (max %d)", d.nv, 4 * kConsumers);
    }
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

// the tree sweeps of solveStep; returns the number of kernels launched and the number of distance slots written
// RN_SWEEP_BATCHED, and what RN_SWEEP_PERSISTENT / RN_SWEEP_CHAIN fall back to when the problem does not fit their shared memory
bool use_batched(const Handle *h) {
    if (h->factor_mode == RN_FACTORS_SHARED) return false;
    if (h->sweep_mode == RN_SWEEP_BATCHED) return true;
    if (h->sweep_mode == RN_SWEEP_PERSISTENT) return !persistent_supported(h);
    if (h->sweep_mode == RN_SWEEP_CHAIN) return !chain_fits(h);
    return false;
}

static rn_status launch_sweeps(Handle *h, cudaStream_t st, bool fuse_prox, int *n_launch, int *n_slots, cudaEvent_t mid = nullptr) {
    const rn_dims &d = h->d;
    if (use_batched(h)) return launch_sweeps_batched(h, st, fuse_prox, n_launch, n_slots, mid);
    const bool chains = h->sweep_mode != RN_SWEEP_PER_STAGE && chain_fits(h);
    const int cs = chains ? h->chain_stage : d.N;   // stages [0, cs) go stage by stage, [cs, N) as chains
    SweepArgs S = make_sweep_args(h, fuse_prox);
    int launches = 0;
    if (chains) {
        const size_t smem = bwd_chain_smem(h);
        RN_CUDA(h, cudaFuncSetAttribute(k_bwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bwd_chain<<<h->h_nps[cs], kChainThreads, smem, st>>>(S, h->cum_dev, cs, d.N - 1);
        launches++;
    }
    for (int s = cs - 1; s >= 0; s--) {
        k_bwd_stage<<<h->h_nps[s], kStageThreads, bwd_stage_smem(h), st>>>(S, h->h_cum[s]);
        launches++;
    }
    if (mid) RN_CUDA(h, cudaEventRecord(mid, st));
    for (int s = 0; s < cs; s++) {
        const int branching = s > 0 && (h->h_nps[s] - h->h_nps[s - 1]) > 0;
        k_fwd_stage<<<h->h_nps[s], kStageThreads, fwd_stage_smem(h), st>>>(S, h->h_cum[s], branching, h->h_cum[s]);
        launches++;
    }
    int slots = h->h_cum[cs];
    if (chains) {
        const size_t smem = fwd_chain_smem(h);
        RN_CUDA(h, cudaFuncSetAttribute(k_fwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fwd_chain<<<h->h_nps[cs], kChainThreads, smem, st>>>(S, h->cum_dev, cs, d.N - 1, h->h_cum[cs]);
        launches++;
        slots += h->h_nps[cs];
    }
    RN_CUDA(h, cudaGetLastError());
    *n_launch = launches;
    *n_slots = slots;
    return RN_OK;
}

static FinalArgs make_final_args(Handle *h) {
    FinalArgs F{};
    F.pri_xi = h->pri_xi; F.pri_psi = h->pri_psi; F.w_xi = h->acc_xi; F.w_psi = h->acc_psi;
    F.dual_xi = h->dual_xi; F.dual_psi = h->dual_psi; F.res_xi = h->res_xi; F.res_psi = h->res_psi;
    F.dist_part = h->dist_part; F.scal = h->scal; F.pinf = h->pinf; F.pinf_part = h->pinf_part;
    F.iter = h->iter_dev; F.done = h->done_ctr;
    F.nodes = h->d.nodes; F.nx = h->d.nx; F.nu = h->d.nu;
    F.step = h->step; F.inv_step = 1 / h->step; F.pen_x = h->pen_x; F.pen_xs = h->pen_xs;
    return F;
}

static int finalize_grid(const Handle *h) { return std::min(h->d.nodes, std::min(2 * h->sm_count, h->pinf_slots)); }

rn_status apg_release_graph(Handle *h) {
    if (h->iter_graph) { cudaGraphExecDestroy(h->iter_graph); h->iter_graph = nullptr; }
    h->graph_sweep = h->graph_factor = -1;
    return RN_OK;
}

// SmpcController::initialiseAlgorithm (:420-450): one memset over the slab that holds the ten dual/primal vectors
rn_status apg_init(Handle *h) {
    RN_CUDA(h, cudaMemsetAsync(h->apg_slab, 0, h->apg_slab_bytes, h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->iter_dev, 0, sizeof(int), h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->done_ctr, 0, sizeof(unsigned int), h->stream));
    // role pointers back to their home buffers
    h->upd_xi = h->yA_xi; h->upd_psi = h->yA_psi; h->xi = h->yB_xi; h->psi = h->yB_psi;
    h->acc_xi = h->wA_xi; h->acc_psi = h->wA_psi;
    return RN_OK;
}

static rn_status ensure_lambda(Handle *h, int iterations) {
    if (iterations + 1 > h->lambda_cap) {
        h->lambda_cap = iterations + 8;
        RN_CHECK(dev_alloc(h, &h->lambda_tab, h->lambda_cap));
        RN_CHECK(dev_alloc(h, &h->pinf, h->lambda_cap));
        h->lambda_ready = 0;
        apg_release_graph(h);
    }
    if (h->lambda_ready < iterations) {
        std::vector<float> tab;
        fill_lambda_table(tab, h->lambda_cap);
        RN_CUDA(h, cudaMemcpyAsync(h->lambda_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        RN_CUDA(h, cudaStreamSynchronize(h->stream));
        h->lambda_ready = h->lambda_cap;
    }
    return RN_OK;
}

// enqueue the kernels of ONE fused iteration on `st`
static rn_status enqueue_iteration(Handle *h, cudaStream_t st, long long *count, cudaEvent_t *ev = nullptr) {
    int nl = 0, slots = 0;
    if (ev) RN_CUDA(h, cudaEventRecord(ev[0], st));
    RN_CHECK(launch_stream(h, st, true, false, h->yA_xi, h->yA_psi, h->yB_xi, h->yB_psi));
    if (ev) RN_CUDA(h, cudaEventRecord(ev[1], st));
    RN_CHECK(launch_sweeps(h, st, true, &nl, &slots, ev ? ev[2] : nullptr));
    if (ev) RN_CUDA(h, cudaEventRecord(ev[3], st));
    FinalArgs F = make_final_args(h);
    F.yA_xi = h->yA_xi; F.yA_psi = h->yA_psi; F.yB_xi = h->yB_xi; F.yB_psi = h->yB_psi;
    F.n_slots = slots;
    F.do_branch = 1; F.do_residual = 1; F.do_update = 1; F.parity_swap = 1; F.log_inf = 1;
    k_finalize<<<finalize_grid(h), kEwThreads, 0, st>>>(F);
    RN_CUDA(h, cudaGetLastError());
    if (ev) RN_CUDA(h, cudaEventRecord(ev[4], st));
    *count = 1 + nl + 1;
    return RN_OK;
}

// one cold-started solve of `iterations` iterations, launched kernel by kernel (no graph) with CUDA events on the
// launching stream around each kernel class; ms_out[c] = mean duration per iteration of class c (rn_prof_class)
rn_status profile_kernels(Handle *h, int iterations, float *ms_out) {
    RN_CHECK(ensure_lambda(h, iterations));
    RN_CHECK(apg_init(h));
    if (use_persistent(h)) {
        // phases of the persistent kernel, clocked by CTA 0 with %globaltimer at every grid barrier of the real run
        cudaEvent_t e0, e1;
        RN_CUDA(h, cudaEventCreate(&e0)); RN_CUDA(h, cudaEventCreate(&e1));
        RN_CUDA(h, cudaEventRecord(e0, h->stream));
        RN_CHECK(enqueue_persistent(h, iterations));
        RN_CUDA(h, cudaEventRecord(e1, h->stream));
        unsigned long long pn[32];
        RN_CUDA(h, cudaMemcpyAsync(pn, h->phase_ns, sizeof(pn), cudaMemcpyDeviceToHost, h->stream));
        h->last_cta_ns.assign(2 * (size_t)h->persist_grid, 0ull);
        RN_CUDA(h, cudaMemcpyAsync(h->last_cta_ns.data(), h->cta_ns, h->last_cta_ns.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        RN_CUDA(h, cudaStreamSynchronize(h->stream));
        float total = 0.f;
        RN_CUDA(h, cudaEventElapsedTime(&total, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        // stream = fused element-wise pass + factor stream + its closing barrier; backward = chains + crown (+ their
        // barriers); forward = crown + chains + prox boxes + the closing barrier.  Indices: cabi.PHASE_NAMES
        unsigned long long ns[4] = {pn[0] + pn[1], 0, 0, 0};
        for (int k = 2; k <= 12; k++) ns[1] += pn[k];
        for (int k = 20; k <= 22; k++) ns[1] += pn[k];
        for (int k = 13; k <= 19; k++) ns[2] += pn[k];
        ns[2] += pn[23] + pn[28] + pn[29];
        memcpy(h->last_phase_ns, pn, sizeof(pn));
        h->last_phase_iters = iterations;
        ms_out[RN_PROF_STREAM] = (float)(ns[0] * 1e-6 / iterations);
        ms_out[RN_PROF_BACKWARD] = (float)(ns[1] * 1e-6 / iterations);
        ms_out[RN_PROF_FORWARD] = (float)((ns[2] + ns[3]) * 1e-6 / iterations);
        const double rest = total - (ns[0] + ns[1] + ns[2] + ns[3]) * 1e-6;
        ms_out[RN_PROF_FINALIZE] = (float)((rest > 0 ? rest : 0) / iterations);   // launch + the one k_finalize
        return RN_OK;
    }
    if (use_batched(h)) RN_CHECK(batched_prepare(h));
    std::vector<cudaEvent_t> ev((size_t)iterations * 5);
    for (auto &e : ev) RN_CUDA(h, cudaEventCreate(&e));
    long long per_iter = 0;
    for (int k = 0; k < iterations; k++) RN_CHECK(enqueue_iteration(h, h->stream, &per_iter, &ev[(size_t)k * 5]));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    double acc[RN_PROF_COUNT_] = {0, 0, 0, 0};
    for (int k = 0; k < iterations; k++)
        for (int c = 0; c < 4; c++) {
            float ms = 0.f;
            RN_CUDA(h, cudaEventElapsedTime(&ms, ev[(size_t)k * 5 + c], ev[(size_t)k * 5 + c + 1]));
            acc[c] += ms;
        }
    for (auto &e : ev) cudaEventDestroy(e);
    for (int c = 0; c < RN_PROF_COUNT_; c++) ms_out[c] = (float)(acc[c] / iterations);
    h->launches += per_iter * iterations;
    h->launches_per_iter = per_iter;
    if (iterations & 1) { std::swap(h->upd_xi, h->xi); std::swap(h->upd_psi, h->psi); }
    return RN_OK;
}

// the persistent cooperative kernel runs iterations 0..iters-1; the last iteration's finalisation (distance branch,
// residual, dual update, infeasibility log) is one k_finalize launch
static rn_status enqueue_persistent(Handle *h, int iterations) {
    RN_CHECK(persistent_prepare(h));
    RN_CUDA(h, cudaMemsetAsync(h->phase_ns, 0, 32 * sizeof(unsigned long long), h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->cta_ns, 0, 2 * 1024 * sizeof(unsigned long long), h->stream));
    if (iterations > 0) {
        RN_CHECK(persistent_launch(h, h->stream, iterations));
        const int last = (iterations - 1) & 1;
        h->acc_xi = last ? h->wB_xi : h->wA_xi; h->acc_psi = last ? h->wB_psi : h->wA_psi;
        FinalArgs F = make_final_args(h);
        F.yA_xi = h->yA_xi; F.yA_psi = h->yA_psi; F.yB_xi = h->yB_xi; F.yB_psi = h->yB_psi;
        F.n_slots = h->dist_world > 1 ? 1 : h->persist_grid;   // several GPUs: the kernel left the global sums in slot 0
        F.pinf4 = h->pinf4;                                    // the last iteration's row of the per-rank log
        F.do_branch = 1; F.do_residual = 1; F.do_update = 1; F.parity_swap = 1; F.log_inf = 1;
        k_finalize<<<finalize_grid(h), kEwThreads, 0, h->stream>>>(F);
        RN_CUDA(h, cudaGetLastError());
        h->launches += 2;
    }
    h->launches_per_iter = 0;
    if (iterations & 1) { std::swap(h->upd_xi, h->xi); std::swap(h->upd_psi, h->psi); }
    if (iterations > 0) h->have_duals = true;
    return RN_OK;
}

bool use_persistent(const Handle *h) { return h->sweep_mode == RN_SWEEP_PERSISTENT && persistent_supported(h); }

rn_status apg_enqueue(Handle *h, int iterations) {
    // opt-in warm start (rn_set_warm_start): from the second solve on, start from the duals the previous solve left
    if (h->warm_start && h->have_duals && iterations > 0 && use_persistent(h)) return apg_warm(h, iterations);
    RN_CHECK(ensure_lambda(h, iterations));
    RN_CHECK(apg_init(h));
    if (use_persistent(h)) return enqueue_persistent(h, iterations);
    if (use_batched(h)) RN_CHECK(batched_prepare(h));
    if (h->factor_mode == RN_FACTORS_SHARED)
        return fail(h, RN_ERR_INVALID, "RN_FACTORS_SHARED needs the persistent sweep (RN_SWEEP_PERSISTENT on a tree it supports)");
    static const bool no_graph = getenv("RN_NO_GRAPH") != nullptr;
    long long per_iter = 0;
    if (no_graph) {
        for (int k = 0; k < iterations; k++) RN_CHECK(enqueue_iteration(h, h->stream, &per_iter));
    } else {
        if (!h->iter_graph || h->graph_sweep != h->sweep_mode || h->graph_factor != h->factor_mode) {
            apg_release_graph(h);
            cudaGraph_t graph = nullptr;
            RN_CUDA(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
            rn_status s = enqueue_iteration(h, h->cap_stream, &per_iter);
            cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
            if (s != RN_OK) { if (graph) cudaGraphDestroy(graph); return s; }
            if (ce != cudaSuccess) return fail(h, RN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&h->iter_graph, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(h, RN_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
            h->graph_sweep = h->sweep_mode; h->graph_factor = h->factor_mode;
            h->launches_per_iter = per_iter;
        }
        per_iter = h->launches_per_iter;
        for (int k = 0; k < iterations; k++) RN_CUDA(h, cudaGraphLaunch(h->iter_graph, h->stream));
    }
    h->launches_per_iter = per_iter;
    h->launches += per_iter * iterations;
    if (iterations & 1) { std::swap(h->upd_xi, h->xi); std::swap(h->upd_psi, h->psi); }   // y_k now lives in the B buffers
    return RN_OK;
}

// `iterations` fused iterations that CONTINUE from the duals in place instead of the cold start of algorithmApg (:1509):
// devVecUpdateXi/Psi is taken as y_k and devVecXi/Psi as y_{k-1}; lambda_host (nullable) replaces the theta recursion for
// these iterations.  The persistent kernel's iteration 0 forms y_0 = w_prev + step (Hx - z), so it is handed
// w_prev = y_k with Hx = z = 0 (y_0 == y_k exactly) and y_{k-1} in the "previous" iterate buffer.
rn_status apg_continue(Handle *h, int iterations, const float *lambda_host) {
    if (!use_persistent(h)) return fail(h, RN_ERR_INVALID, "rn_apg_continue needs the persistent sweep on a tree it supports");
    if (iterations <= 0) return RN_OK;
    const rn_dims &d = h->d;
    RN_CHECK(ensure_lambda(h, iterations));
    if (lambda_host) {
        RN_CUDA(h, cudaMemcpyAsync(h->lambda_tab, lambda_host, (size_t)iterations * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        RN_CUDA(h, cudaStreamSynchronize(h->stream));   // the caller's array may be pageable
        h->lambda_ready = 0;                            // the next cold solve restores the theta recursion
    }
    const size_t bx = (size_t)d.nodes * 2 * d.nx * sizeof(float), bp = (size_t)d.nodes * d.nu * sizeof(float);
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    // y_k -> wB first (it never aliases a y buffer), then y_{k-1} -> yB unless it already lives there
    RN_CUDA(h, cudaMemcpyAsync(h->wB_xi, h->upd_xi, bx, k, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->wB_psi, h->upd_psi, bp, k, h->stream));
    if (h->xi != h->yB_xi) {
        RN_CUDA(h, cudaMemcpyAsync(h->yB_xi, h->xi, bx, k, h->stream));
        RN_CUDA(h, cudaMemcpyAsync(h->yB_psi, h->psi, bp, k, h->stream));
    }
    RN_CUDA(h, cudaMemsetAsync(h->pri_xi, 0, bx, h->stream)); RN_CUDA(h, cudaMemsetAsync(h->pri_psi, 0, bp, h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->dual_xi, 0, bx, h->stream)); RN_CUDA(h, cudaMemsetAsync(h->dual_psi, 0, bp, h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->iter_dev, 0, sizeof(int), h->stream));
    RN_CUDA(h, cudaMemsetAsync(h->done_ctr, 0, sizeof(unsigned int), h->stream));
    h->upd_xi = h->yA_xi; h->upd_psi = h->yA_psi; h->xi = h->yB_xi; h->psi = h->yB_psi;
    h->acc_xi = h->wA_xi; h->acc_psi = h->wA_psi;
    return enqueue_persistent(h, iterations);
}

// warm start (opt-in; the reference always cold-starts, :420-450, :1509): y_0 = y_{-1} = the duals the previous
// solve left in devVecUpdateXi/Psi, theta restarted at 1
rn_status apg_warm(Handle *h, int iterations) {
    if (!use_persistent(h)) return fail(h, RN_ERR_INVALID, "warm start needs the persistent sweep on a tree it supports");
    const rn_dims &d = h->d;
    const size_t bx = (size_t)d.nodes * 2 * d.nx * sizeof(float), bp = (size_t)d.nodes * d.nu * sizeof(float);
    RN_CUDA(h, cudaMemcpyAsync(h->xi, h->upd_xi, bx, cudaMemcpyDeviceToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->psi, h->upd_psi, bp, cudaMemcpyDeviceToDevice, h->stream));
    return apg_continue(h, iterations, nullptr);
}

rn_status apg_step(Handle *h, rn_step_kind kind, float lambda) {
    const rn_dims &d = h->d;
    const size_t n_xi = (size_t)d.nodes * 2 * d.nx, n_psi = (size_t)d.nodes * d.nu;
    switch (kind) {
        case RN_STEP_EXTRAPOLATE: {
            const int grid = (int)std::min<size_t>((n_xi + n_psi + 255) / 256, (size_t)h->sm_count * 8);
            k_extrapolate<<<grid, 256, 0, h->stream>>>(n_xi, n_psi, lambda, h->upd_xi, h->upd_psi, h->xi, h->psi, h->acc_xi, h->acc_psi);
            h->launches += 1;
            break;
        }
        case RN_STEP_SOLVE: {
            int nl = 0, slots = 0;
            if (use_batched(h)) RN_CHECK(batched_prepare(h));
            RN_CHECK(launch_stream(h, h->stream, false, false, h->upd_xi, h->upd_psi, h->xi, h->psi));
            RN_CHECK(launch_sweeps(h, h->stream, false, &nl, &slots));
            h->launches += 1 + nl;
            break;
        }
        case RN_STEP_PROX: {
            SweepArgs S = make_sweep_args(h, true);
            k_prox_boxes<<<d.nodes, kEwThreads, 0, h->stream>>>(S, d.nodes);
            FinalArgs F = make_final_args(h);
            F.n_slots = d.nodes; F.do_branch = 1;
            k_finalize<<<finalize_grid(h), kEwThreads, 0, h->stream>>>(F);
            h->launches += 2;
            break;
        }
        case RN_STEP_RESIDUAL: {
            FinalArgs F = make_final_args(h);
            F.do_residual = 1;
            k_finalize<<<finalize_grid(h), kEwThreads, 0, h->stream>>>(F);
            h->launches += 1;
            break;
        }
        case RN_STEP_DUAL_UPDATE: {
            FinalArgs F = make_final_args(h);
            F.do_update = 1; F.yA_xi = h->upd_xi; F.yA_psi = h->upd_psi;
            k_finalize<<<finalize_grid(h), kEwThreads, 0, h->stream>>>(F);
            h->launches += 1;
            break;
        }
        default: return fail(h, RN_ERR_INVALID, "rn_step: unknown step %d", (int)kind);
    }
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

rn_status profile_stream(Handle *h, int reps, float *mean_ms) {
    cudaEvent_t e0, e1;
    RN_CUDA(h, cudaEventCreate(&e0)); RN_CUDA(h, cudaEventCreate(&e1));
    RN_CHECK(launch_stream(h, h->stream, true, true, h->yA_xi, h->yA_psi, h->yB_xi, h->yB_psi));   // warm-up
    RN_CUDA(h, cudaEventRecord(e0, h->stream));
    for (int k = 0; k < reps; k++) RN_CHECK(launch_stream(h, h->stream, true, true, h->yA_xi, h->yA_psi, h->yB_xi, h->yB_psi));
    RN_CUDA(h, cudaEventRecord(e1, h->stream));
    RN_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    RN_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->launches += reps + 1;
    *mean_ms = ms / reps;
    h->last_stream_ms = *mean_ms;
    return RN_OK;
}

rn_status clamp_control(Handle *h) {   // projectionBox<<<1,nu>>> with the node-0 preconditioned bounds (:1649)
    k_clamp_vec<<<ceil_div(h->d.nu, 128), 128, 0, h->stream>>>(h->d.nu, h->control_action, h->sumin, h->sumax);
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

rn_status move_forward(Handle *h) {
    k_move_forward<<<ceil_div(h->d.nx, 128), 128, 0, h->stream>>>(h->d.nx, h->d.nu, h->xcur, h->B, h->control_action, h->e, h->X,
                                                                 h->state_update);
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

}  // namespace rn
