// rn_persist.cu -- the whole APG loop of SmpcController::algorithmApg as ONE persistent cooperative kernel.
//
// Reference hot loop: /root/reference/src/SmpcController.cu:1500-1525 (about 430 launches per iteration).
// Here every iteration runs inside one resident grid (one CTA per SM, 16 warps) with
// software grid barriers between the phases ("persistent kernel"):
//
//   phase S  factor stream.  Work unit = (node, matrix in {D, F, Phi, Psi}).  The loader warp keeps a 6-deep ring of
//            16 KB stages full with 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP) of the node's packed Engine
//            factor matrix and fetches the unit's dual vectors with cp.async into a 2-deep vector ring, one unit
//            ahead; the ring runs across the barriers, so the next iteration's first matrices arrive while the sweeps
//            run.  The consumer prologue is the fused element-wise pass: finalisation of the PREVIOUS iteration's prox
//            (distance branch), fixed-point residual, dual update y+ = w + step*res, infeasibility log, and the Nesterov
//            extrapolation w = (1+l) y+ - l y of THIS iteration (:535-557, :792-864, :1480-1496) -- the duals are read
//            once and written once per iteration.  Then partial products D xi_w, F psi_w, Phi xi_w, Psi psi_w.
//   phase B  backward sweep (:593-673).  Below the last branching stage every scenario is an independent chain owned
//            by one CTA: the stage recursion is split into scans (q = c + q_child; sigma = beta + r_child,
//            r = sigma + D xi + F psi + G q_child) and true GEMMs across the chain's stages against the shared
//            matrices (G, Omega_chain, Theta_chain, L) -- no barrier per stage.  Above it ("crown") one grid barrier per
//            stage with the child->parent sums fused into the node's GEMVs (solveSumChildren, Utilities.cu:168-201).
//   phase F  forward sweep (:675-747): crown stage by stage, then per chain u-scan, B U GEMM, x-scan, and the
//            epilogue Hx = sysF x, Hu = sysG u, t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial
//            sums of the two global distances of proximalFunG (:792, :810).
//
// The last iteration's finalisation is done by k_finalize (rn_apg.cu) after the kernel.
#include <algorithm>

#include "rn_internal.h"
#include "rn_device.cuh"

namespace rn {

constexpr int kPC = 512;                        // threads per CTA: 16 warps -> 128 registers per thread
constexpr int kPT = kPC;                        // phase S roles: 11 GEMV warps, 4 element-wise warps, 1 loader warp
constexpr int kPStages = 6;
constexpr int kPStageFloats = 4096;
constexpr int kPStageStride = kPStageFloats + 32;
constexpr int kGemvWarps = 11;                  // phase S: warps that multiply the streamed matrices
constexpr int kEwWarps = 4;                     // phase S: warps that run the fused element-wise pass
constexpr int kLoaderWarp = kGemvWarps + kEwWarps;   // phase S: the warp that drives the TMA ring and the vector ring
static_assert(kLoaderWarp == kPC / 32 - 1, "role split must cover the CTA");
constexpr int kVecSlots = 3;
constexpr int kVecCount = 9;                    // Hx, w_prev, z, y_prev for xi and for psi, diag
constexpr int kVStride = 128;                   // floats per staged vector (>= max(2nx, nu))
constexpr int kWStride = 2 * kVStride;          // w of one node: xi part | psi part
constexpr int kTMax = 24;                       // longest chain (stages below the last branching stage)
constexpr int kCG = 6;                          // columns per thread in the chain GEMMs (4 groups x 6)
constexpr int kTP = 24;                         // padded column count of the transposed right-hand sides
constexpr int kDimMax = 128;                    // max(nx, nu, nv) supported by this kernel

struct PArgs {
    const int *parent, *child_first, *child_count, *omega_idx, *cum;
    int N, cs, K, nodes, n_mats, df_mode, iters, nx, nu, nv, cols_per_chunk;
    const float *mat[4];                        // D, F, Phi, Psi (packed per node, Engine.cu:201-207)
    const float *Omega, *Theta, *G, *L, *B, *diag;
    const float *beta, *uhat, *e, *xcur, *uprev, *uhat_prev, *sxmin, *sxmax, *sxs, *sumin, *sumax;
    float *Yxi[2], *Ypsi[2], *Wxi[2], *Wpsi[2];
    float *pri_xi, *pri_psi, *dual_xi, *dual_psi;
    float *part[4];                             // D xi_w, F psi_w, Phi xi_w, Psi psi_w   [nodes*nv] each
    float *c, *q, *r, *sigma, *V, *U, *X, *LV;
    double *dist_part;                          // [2*grid]
    float *pinf, *pinf_part;                    // [iters], [grid*6]
    const float *lambda_tab;
    unsigned int *bar;
    int *iter_dev;
    unsigned long long *phase_ns;               // [32] fine-grained phase clock of CTA 0 (see kPhaseNames)
    float step, inv_step, pen_x, pen_xs;
};

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kPC) : "memory"); }
__device__ __forceinline__ void ewbar() { asm volatile("bar.sync 2, %0;" ::"n"(kEwWarps * 32) : "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// fine-grained clock of CTA 0 (thread 0): accumulates the time since the previous stamp into phase_ns[idx]
__device__ unsigned long long g_t_prev;
__device__ __forceinline__ void dstamp(unsigned long long *phase_ns, int idx) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long now = globaltimer();
        phase_ns[idx] += now - g_t_prev;
        g_t_prev = now;
    }
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid barrier over the consumer threads of all CTAs (the grid is co-resident: cooperative launch).  Same protocol as
// cooperative_groups::grid_group::sync: CTA barrier, one thread fences + arrives + spins + fences, CTA barrier.
__device__ __forceinline__ void grid_sync(unsigned int *ctr, unsigned int &target) {
    cbar();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) {}
        __threadfence();
    }
    cbar();
}

// ys[r] = sum_c A[r + c*lda] xs[c]  (A global, xs/ys/scratch shared; 512 consumer threads; ends with a barrier)
__device__ __forceinline__ void cgemv(const float *__restrict__ A, int lda, int m, int n, const float *xs, float *ys,
                                      float *scratch) {
    const int t = threadIdx.x;
    int slots = (m + 31) & ~31;
    const int G = kPC / slots;
    const int g = t / slots, rr = t - g * slots;
    if (g < G && rr < m) {
        float acc = 0.f;
#pragma unroll 16
        for (int c = g; c < n; c += G) acc = fmaf(__ldg(A + rr + (size_t)c * lda), xs[c], acc);
        scratch[g * m + rr] = acc;
    }
    cbar();
    if (t < m) {
        float s = scratch[t];
        for (int gg = 1; gg < G; gg++) s += scratch[gg * m + t];
        ys[t] = s;
    }
    cbar();
}

// shared-memory layout (float offsets from the dynamic shared-memory base; all compile-time constants so that every
// access is an LDS/STS with an immediate offset even inside the non-inlined role functions)
constexpr int kOffRing = 0;
constexpr int kOffVec = kOffRing + kPStages * kPStageStride;            // kVecSlots x kVecCount x kVStride
constexpr int kOffW = kOffVec + kVecSlots * kVecCount * kVStride;       // 2 x kWStride  (xi part | psi part)
constexpr int kOffRed = kOffW + 2 * kWStride;                           // 2 x kGemvWarps x kDimMax
constexpr int kOffTail = kOffRed + 2 * kGemvWarps * kDimMax;            // 5 x [kDimMax][kTP]
constexpr int kOffScr = kOffTail + 5 * kDimMax * kTP;                   // kPC floats
constexpr int kOffDsh = kOffScr + kPC;                                  // 2 x 16 doubles
constexpr int kOffCsh = kOffDsh + 2 * 2 * (kPC / 32);                   // 2 x 16 candidates (3 words each)
constexpr int kOffSd = kOffCsh + 3 * 2 * (kPC / 32);                    // d1, d2
constexpr int kOffBar = kOffSd + 4;                                     // mbarriers
constexpr int kNumBars = 2 * kPStages + 2 * kVecSlots + 8;
constexpr int kSmemFloats = kOffBar + 2 * kNumBars;
static_assert(kOffDsh % 2 == 0 && kOffBar % 2 == 0, "8-byte alignment of the double / mbarrier areas");

__device__ __forceinline__ float *smem_f(int off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    return reinterpret_cast<float *>(smem_raw) + off;
}

struct TailSmem {
    float *A0, *A1, *A2, *A3;   // four [kDimMax][kTP] arrays
    float *scr;                 // kPC floats
};
__device__ __forceinline__ TailSmem tail_smem() {
    TailSmem S;
    S.A0 = smem_f(kOffTail); S.A1 = S.A0 + kDimMax * kTP; S.A2 = S.A1 + kDimMax * kTP; S.A3 = S.A2 + kDimMax * kTP;
    S.scr = smem_f(kOffScr);
    return S;
}

// Hx, Hu and the box part of proximalFunG for element `el` of node i (x, u already known); returns through s1/s2 the
// squared distance contributions.
__device__ __forceinline__ void prox_element(const PArgs &P, int i, int el, float xv_or_uv, const float *wxi, const float *wpsi,
                                             double &s1, double &s2) {
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu;
    const float dgv = __ldg(P.diag + (size_t)i * ny + el);
    if (el < 2 * nx) {
        const int j = el < nx ? el : el - nx;
        const size_t k = (size_t)i * 2 * nx + el, kb = (size_t)i * nx + j;
        const float hx = dgv * xv_or_uv;
        P.pri_xi[k] = hx;
        const float tt = hx + P.inv_step * __ldcg(wxi + k);
        const float z = el < nx ? clampf(tt, __ldg(P.sxmin + kb), __ldg(P.sxmax + kb))
                                : clampf(tt, __ldg(P.sxs + kb), __int_as_float(0x7F7F7F7F));
        P.dual_xi[k] = z;
        const float df = tt + -1.f * z;
        if (el < nx) s1 += (double)df * df; else s2 += (double)df * df;
    } else {
        const int j = el - 2 * nx;
        const size_t k = (size_t)i * nu + j;
        const float hu = dgv * xv_or_uv;
        P.pri_psi[k] = hu;
        P.dual_psi[k] = clampf(hu + P.inv_step * __ldcg(wpsi + k), __ldg(P.sumin + k), __ldg(P.sumax + k));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// chains (stages cs .. N-1 of scenario j).  Every step is a small non-inlined function so that none of them is
// register-critical (ptxas otherwise schedules the WHOLE kernel, including the stream loop, for minimum registers).
// Shared arrays A0..A4 are [element][stage] (kTP stages per row).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chain_node(const PArgs &P, int s, int j) { return __ldg(P.cum + P.cs + s) + j; }

// dst[e][s] = src[node(s)*dim + e] for all stages of the chain (coalesced over e, stages 4 apart per thread)
__device__ __noinline__ void chain_stage_in(const PArgs &P, int j, const float *src, int dim, int dst_off, bool coherent) {
    float *dst = smem_f(dst_off);
    const int t = threadIdx.x, e = t & (kDimMax - 1), s0 = t >> 7, T = P.N - P.cs;
    if (e >= dim) return;
    float v[kTMax / 4];
#pragma unroll
    for (int k = 0; k < kTMax / 4; k++) {
        const int s = s0 + 4 * k;
        if (s < T) {
            const float *ptr = src + (size_t)chain_node(P, s, j) * dim + e;
            v[k] = coherent ? __ldcg(ptr) : __ldg(ptr);
        }
    }
#pragma unroll
    for (int k = 0; k < kTMax / 4; k++) {
        const int s = s0 + 4 * k;
        if (s < T) dst[e * kTP + s] = v[k];
    }
}

// dst[node(s)*dim + e] = src[e][s]
__device__ __noinline__ void chain_stage_out(const PArgs &P, int j, float *dst, int dim, int src_off) {
    const float *src = smem_f(src_off);
    const int t = threadIdx.x, e = t & (kDimMax - 1), s0 = t >> 7, T = P.N - P.cs;
    if (e >= dim) return;
#pragma unroll
    for (int k = 0; k < kTMax / 4; k++) {
        const int s = s0 + 4 * k;
        if (s < T) dst[(size_t)chain_node(P, s, j) * dim + e] = src[e * kTP + s];
    }
}

// Y[r][s] = sum_k M[r + k*m] X[k][s]   (M: m x kdim, global and constant; X, Y shared; thread = (row, 6 stages))
__device__ __noinline__ void chain_gemm(const float *__restrict__ M, int m, int kdim, int x_off, int y_off) {
    const float *Xt = smem_f(x_off);
    float *Y = smem_f(y_off);
    const int t = threadIdx.x, row = t & (kDimMax - 1), col0 = (t >> 7) * kCG;
    if (row >= m) return;
    float acc[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float *mp = M + row;
    const float *xp = Xt + col0;
#pragma unroll 8
    for (int k = 0; k < kdim; k++) {
        const float mv = __ldg(mp + (size_t)k * m);
        const float2 x0 = *reinterpret_cast<const float2 *>(xp + k * kTP);
        const float2 x1 = *reinterpret_cast<const float2 *>(xp + k * kTP + 2);
        const float2 x2 = *reinterpret_cast<const float2 *>(xp + k * kTP + 4);
        acc[0] = fmaf(mv, x0.x, acc[0]); acc[1] = fmaf(mv, x0.y, acc[1]);
        acc[2] = fmaf(mv, x1.x, acc[2]); acc[3] = fmaf(mv, x1.y, acc[3]);
        acc[4] = fmaf(mv, x2.x, acc[4]); acc[5] = fmaf(mv, x2.y, acc[5]);
    }
#pragma unroll
    for (int c = 0; c < kCG; c++) Y[row * kTP + col0 + c] = acc[c];
}

// q-scan in place on A0: A0 holds c on entry, q_bar (= q of the child, 0 at the leaf) on exit; q = c + q_bar (:651-658)
__device__ __noinline__ void chain_qscan(const PArgs &P, int j) {
    float *Qb = smem_f(kOffTail);
    const int t = threadIdx.x, T = P.N - P.cs;
    if (t >= P.nx) return;
    float qrun = 0.f;
    for (int s = T - 1; s >= 0; s--) { const float cv = Qb[t * kTP + s]; Qb[t * kTP + s] = qrun; qrun = cv + qrun; }
    P.q[(size_t)chain_node(P, 0, j) * P.nx + t] = qrun;   // head of the chain, for the crown
}

// r-scan: A2 = beta -> sigma (df: r), A3 = D xi, A4 = F psi, A1 = G q_bar.
// sigma = beta + r_child (:599); r = ((sigma + D xi) + F psi) + G q_bar (:631-646)
__device__ __noinline__ void chain_rscan(const PArgs &P, int j) {
    float *Y3 = smem_f(kOffTail + kDimMax * kTP), *Sg = Y3 + kDimMax * kTP, *Ad = Sg + kDimMax * kTP, *Af = Ad + kDimMax * kTP;
    const int t = threadIdx.x, T = P.N - P.cs, nv = P.nv;
    if (t >= nv) return;
    float rrun = 0.f;
    for (int s = T - 1; s >= 0; s--) {
        const float sg = Sg[t * kTP + s] + rrun;
        rrun = ((sg + Ad[t * kTP + s]) + Af[t * kTP + s]) + Y3[t * kTP + s];
        Sg[t * kTP + s] = P.df_mode ? rrun : sg;
        Ad[t * kTP + s] = sg;   // sigma, written out by chain_stage_out
    }
    P.r[(size_t)chain_node(P, 0, j) * nv + t] = rrun;
}

// v = ((-1/2 Omega sigma + Theta q_bar) + Psi psi) + Phi xi (:604-627) [df: v = -1/2 Omega r]; A1 = Omega x, A3 = Theta q_bar
// on entry, A1 = v on exit (and V in global memory)
__device__ __noinline__ void chain_vcombine(const PArgs &P, int j) {
    float *Y1 = smem_f(kOffTail + kDimMax * kTP), *Y2 = Y1 + 2 * kDimMax * kTP;
    const int t = threadIdx.x, e = t & (kDimMax - 1), s0 = t >> 7, T = P.N - P.cs, nv = P.nv;
    if (e >= nv) return;
    float b3[kTMax / 4], b2[kTMax / 4];
    if (!P.df_mode) {
#pragma unroll
        for (int k = 0; k < kTMax / 4; k++) {
            const int s = s0 + 4 * k;
            if (s < T) {
                const size_t idx = (size_t)chain_node(P, s, j) * nv + e;
                b3[k] = __ldcg(P.part[3] + idx); b2[k] = __ldcg(P.part[2] + idx);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kTMax / 4; k++) {
        const int s = s0 + 4 * k;
        if (s < T) {
            float v;
            if (P.df_mode) v = -0.5f * Y1[e * kTP + s];
            else v = ((-0.5f * Y1[e * kTP + s] + Y2[e * kTP + s]) + b3[k]) + b2[k];
            Y1[e * kTP + s] = v;
            P.V[(size_t)chain_node(P, s, j) * nv + e] = v;
        }
    }
}

__device__ __noinline__ void tail_backward(const PArgs &P, int j) {
    constexpr int A0 = kOffTail, A1 = A0 + kDimMax * kTP, A2 = A1 + kDimMax * kTP, A3 = A2 + kDimMax * kTP, A4 = A3 + kDimMax * kTP;
    const int nx = P.nx, nv = P.nv, nu = P.nu;
    chain_stage_in(P, j, P.c, nx, A0, true);
    chain_stage_in(P, j, P.beta, nv, A2, false);
    chain_stage_in(P, j, P.part[0], nv, A3, true);
    chain_stage_in(P, j, P.part[1], nv, A4, true);
    cbar();
    dstamp(P.phase_ns, 3);
    chain_qscan(P, j);
    cbar();
    dstamp(P.phase_ns, 4);
    chain_gemm(P.G, nv, nx, A0, A1);                       // A1 = G q_bar
    cbar();
    dstamp(P.phase_ns, 5);
    chain_rscan(P, j);                                      // A2 = sigma (df: r), A3 = sigma
    cbar();
    dstamp(P.phase_ns, 6);
    const int oi = __ldg(P.omega_idx + chain_node(P, 0, j));   // one Omega/Theta per chain (Engine.cu:210-221)
    chain_stage_out(P, j, P.sigma, nv, A3);
    cbar();
    chain_gemm(P.Omega + (size_t)oi * nv * nv, nv, nv, A2, A1);            // A1 = Omega sigma
    if (!P.df_mode) chain_gemm(P.Theta + (size_t)oi * nv * nx, nv, nx, A0, A3);   // A3 = Theta q_bar
    cbar();
    chain_vcombine(P, j);                                   // A1 = v
    cbar();
    dstamp(P.phase_ns, 7);
    chain_gemm(P.L, nu, nv, A1, A2);                        // A2 = L v   (:701, :727 -- forward sweep of the reference)
    cbar();
    chain_stage_out(P, j, P.LV, nu, A2);
    cbar();
    dstamp(P.phase_ns, 8);
}

// u-scan: A2 = uhat, A3 = L v -> A0 = u.  u = ((uhat + u_par) - uhat_par) + L v (:722-728); the chain's first stage is a
// branching stage of the reference's loop when it has more nodes than its parent stage (:699-719): same sums, other
// association
__device__ __noinline__ void chain_uscan(const PArgs &P, int j, bool head_br) {
    float *Ut = smem_f(kOffTail), *Uh = Ut + 2 * kDimMax * kTP, *Lv = Uh + kDimMax * kTP;
    const int t = threadIdx.x, T = P.N - P.cs, nu = P.nu;
    if (t >= nu) return;
    const int par0 = __ldg(P.parent + chain_node(P, 0, j));
    float up = par0 < 0 ? __ldg(P.uprev + t) : __ldcg(P.U + (size_t)par0 * nu + t);
    float uhp = par0 < 0 ? __ldg(P.uhat_prev + t) : __ldg(P.uhat + (size_t)par0 * nu + t);
    for (int s = 0; s < T; s++) {
        const float uh = Uh[t * kTP + s], lv = Lv[t * kTP + s];
        const float u = (s == 0 && head_br) ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;
        Ut[t * kTP + s] = u;
        up = u; uhp = uh;
    }
}

// x-scan in place on A1 (= B u on entry, x on exit), A4 = e.  x = (x_par + e) + B u (:730-737)
__device__ __noinline__ void chain_xscan(const PArgs &P, int j, bool head_br) {
    float *Xt = smem_f(kOffTail + kDimMax * kTP), *Ev = Xt + 3 * kDimMax * kTP;
    const int t = threadIdx.x, T = P.N - P.cs, nx = P.nx;
    if (t >= nx) return;
    const int par0 = __ldg(P.parent + chain_node(P, 0, j));
    float xrun = par0 < 0 ? __ldg(P.xcur + t) : __ldcg(P.X + (size_t)par0 * nx + t);
    for (int s = 0; s < T; s++) {
        const float ev = Ev[t * kTP + s], bu = Xt[t * kTP + s];
        const float x = (s == 0 && head_br) ? xrun + (ev + bu) : (xrun + ev) + bu;
        Xt[t * kTP + s] = x;
        xrun = x;
    }
}

// Hx = sysF x, Hu = sysG u (:744-747), t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial sums of
// the two global distances (:792, :810) for every node of the chain.  A0 = u, A1 = x.  thread = element, 4 stages a batch
__device__ __noinline__ void chain_epilogue(const PArgs &P, int j, const float *wxi, const float *wpsi, double &s1, double &s2) {
    const float *Ut = smem_f(kOffTail), *Xt = Ut + kDimMax * kTP;
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu, T = P.N - P.cs;
    double l1 = 0, l2 = 0;
    for (int el = threadIdx.x; el < ny; el += kPC) {
        const bool isx = el < 2 * nx;
        const int jx = el < nx ? el : el - nx, ju = el - 2 * nx;
        const float *src = isx ? Xt + jx * kTP : Ut + ju * kTP;
        for (int sb = 0; sb < T; sb += 4) {
            float dgv[4], wv[4], lo[4], hi[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int s = sb + b;
                if (s < T) {
                    const size_t i = (size_t)chain_node(P, s, j);
                    dgv[b] = __ldg(P.diag + i * ny + el);
                    if (isx) {
                        const size_t kb = i * nx + jx;
                        wv[b] = __ldcg(wxi + i * 2 * nx + el);
                        lo[b] = el < nx ? __ldg(P.sxmin + kb) : __ldg(P.sxs + kb);
                        hi[b] = el < nx ? __ldg(P.sxmax + kb) : __int_as_float(0x7F7F7F7F);
                    } else {
                        const size_t kk = i * nu + ju;
                        wv[b] = __ldcg(wpsi + kk);
                        lo[b] = __ldg(P.sumin + kk);
                        hi[b] = __ldg(P.sumax + kk);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int s = sb + b;
                if (s < T) {
                    const size_t i = (size_t)chain_node(P, s, j);
                    const float h = dgv[b] * src[s];
                    const float tt = h + P.inv_step * wv[b];
                    const float z = clampf(tt, lo[b], hi[b]);
                    if (isx) {
                        P.pri_xi[i * 2 * nx + el] = h; P.dual_xi[i * 2 * nx + el] = z;
                        const float df = tt + -1.f * z;
                        if (el < nx) l1 += (double)df * df; else l2 += (double)df * df;
                    } else { P.pri_psi[i * nu + ju] = h; P.dual_psi[i * nu + ju] = z; }
                }
            }
        }
    }
    s1 += l1; s2 += l2;
}

// chains: forward.  wxi/wpsi = this iteration's accelerated duals.
__device__ __noinline__ void tail_forward(const PArgs &P, int j, const float *wxi, const float *wpsi, double &s1, double &s2) {
    constexpr int A0 = kOffTail, A1 = A0 + kDimMax * kTP, A2 = A1 + kDimMax * kTP, A3 = A2 + kDimMax * kTP, A4 = A3 + kDimMax * kTP;
    const int nx = P.nx, nu = P.nu;
    const bool head_br = P.cs > 0 && (__ldg(P.cum + P.cs + 1) - __ldg(P.cum + P.cs)) > (__ldg(P.cum + P.cs) - __ldg(P.cum + P.cs - 1));
    chain_stage_in(P, j, P.uhat, nu, A2, false);
    chain_stage_in(P, j, P.LV, nu, A3, true);
    chain_stage_in(P, j, P.e, nx, A4, false);
    cbar();
    dstamp(P.phase_ns, 15);
    chain_uscan(P, j, head_br);                             // A0 = u
    cbar();
    dstamp(P.phase_ns, 16);
    chain_stage_out(P, j, P.U, nu, A0);
    chain_gemm(P.B, nx, nu, A0, A1);                        // A1 = B u   (:715, :736)
    cbar();
    dstamp(P.phase_ns, 17);
    chain_xscan(P, j, head_br);                             // A1 = x
    cbar();
    dstamp(P.phase_ns, 18);
    chain_stage_out(P, j, P.X, nx, A1);
    chain_epilogue(P, j, wxi, wpsi, s1, s2);
    cbar();
    dstamp(P.phase_ns, 19);
}

// ---------------------------------------------------------------------------------------------------------------
// crown (stages above the chains): one node per CTA and per call
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void crown_backward(const PArgs &P, int i) {
    const TailSmem S = tail_smem();
    const int nx = P.nx, nv = P.nv, nu = P.nu, t = threadIdx.x;
    float *qs = S.A0, *sg = qs + kDimMax, *y1 = sg + kDimMax, *y2 = y1 + kDimMax, *y3 = y2 + kDimMax, *vs = y3 + kDimMax;
    const int c0 = __ldg(P.child_first + i), nc = __ldg(P.child_count + i);
    if (t < nx) {   // solveSumChildren (Utilities.cu:168-201)
        float s = 0.f;
        if (nc > 0) { s = __ldcg(P.q + (size_t)c0 * nx + t); for (int c = 1; c < nc; c++) s += __ldcg(P.q + (size_t)(c0 + c) * nx + t); }
        qs[t] = s;
    }
    if (t >= kDimMax && t < kDimMax + nv) {
        const int k = t - kDimMax;
        float s = __ldg(P.beta + (size_t)i * nv + k);
        if (nc > 0) {
            float rs = __ldcg(P.r + (size_t)c0 * nv + k);
            for (int c = 1; c < nc; c++) rs += __ldcg(P.r + (size_t)(c0 + c) * nv + k);
            s += rs;
        }
        sg[k] = s;
        P.sigma[(size_t)i * nv + k] = s;
    }
    cbar();
    cgemv(P.G, nv, nv, nx, qs, y3, S.scr);
    if (t < nv) {
        const size_t k = (size_t)i * nv + t;
        const float rr = ((sg[t] + __ldcg(P.part[0] + k)) + __ldcg(P.part[1] + k)) + y3[t];
        y3[t] = rr;
        P.r[k] = rr;
    }
    if (t >= kDimMax && t < kDimMax + nx) {
        const int k = t - kDimMax;
        P.q[(size_t)i * nx + k] = __ldcg(P.c + (size_t)i * nx + k) + qs[k];
    }
    cbar();
    const int oi = __ldg(P.omega_idx + i);
    const float *Om = P.Omega + (size_t)oi * nv * nv, *Th = P.Theta + (size_t)oi * nv * nx;
    if (P.df_mode) {
        cgemv(Om, nv, nv, nv, y3, y1, S.scr);
        if (t < nv) { const float v = -0.5f * y1[t]; vs[t] = v; P.V[(size_t)i * nv + t] = v; }
    } else {
        cgemv(Om, nv, nv, nv, sg, y1, S.scr);
        cgemv(Th, nv, nv, nx, qs, y2, S.scr);
        if (t < nv) {
            const size_t k = (size_t)i * nv + t;
            const float v = ((-0.5f * y1[t] + y2[t]) + __ldcg(P.part[3] + k)) + __ldcg(P.part[2] + k);
            vs[t] = v;
            P.V[k] = v;
        }
    }
    cbar();
    cgemv(P.L, nu, nu, nv, vs, y1, S.scr);
    if (t < nu) P.LV[(size_t)i * nu + t] = y1[t];
    cbar();
}

__device__ __noinline__ void crown_forward(const PArgs &P, int i, int branching, const float *wxi,
                                              const float *wpsi, double &s1, double &s2) {
    const TailSmem S = tail_smem();
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu, t = threadIdx.x;
    float *us = S.A0, *xs = us + kDimMax, *bu = xs + kDimMax;
    const int par = __ldg(P.parent + i);
    if (t < nu) {
        const float uh = __ldg(P.uhat + (size_t)i * nu + t);
        const float up = par < 0 ? __ldg(P.uprev + t) : __ldcg(P.U + (size_t)par * nu + t);
        const float uhp = par < 0 ? __ldg(P.uhat_prev + t) : __ldg(P.uhat + (size_t)par * nu + t);
        const float lv = __ldcg(P.LV + (size_t)i * nu + t);
        float u;
        if (branching) u = (up + -1.f * uhp) + (uh + lv);          // :701-710
        else u = ((uh + up) + -1.f * uhp) + lv;                    // :683-693, :722-728
        us[t] = u;
        P.U[(size_t)i * nu + t] = u;
    }
    cbar();
    cgemv(P.B, nx, nx, nu, us, bu, S.scr);
    if (t < nx) {
        const float xp = par < 0 ? __ldg(P.xcur + t) : __ldcg(P.X + (size_t)par * nx + t);
        const float ei = __ldg(P.e + (size_t)i * nx + t);
        const float x = branching ? xp + (ei + bu[t]) : (xp + ei) + bu[t];   // :712-719 / :730-737
        xs[t] = x;
        P.X[(size_t)i * nx + t] = x;
    }
    cbar();
    for (int el = t; el < ny; el += kPC) {
        const float v = el < 2 * nx ? xs[el < nx ? el : el - nx] : us[el - 2 * nx];
        prox_element(P, i, el, v, wxi, wpsi, s1, s2);
    }
    cbar();
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
// shared-memory pipeline of phase S
struct Pipe {
    float *ring, *vec, *wbuf, *red;
    uint64_t *full, *empty, *vfull, *vempty, *wfull, *wempty, *rfull, *rempty;
};
__device__ __forceinline__ Pipe pipe_smem() {
    Pipe M;
    M.ring = smem_f(kOffRing); M.vec = smem_f(kOffVec); M.wbuf = smem_f(kOffW); M.red = smem_f(kOffRed);
    M.full = reinterpret_cast<uint64_t *>(smem_f(kOffBar));
    M.empty = M.full + kPStages; M.vfull = M.empty + kPStages; M.vempty = M.vfull + kVecSlots;
    M.wfull = M.vempty + kVecSlots; M.wempty = M.wfull + 2; M.rfull = M.wempty + 2; M.rempty = M.rfull + 2;
    return M;
}
struct Slice { int u_begin, u_end, node_first, node_last; };   // this CTA's share of the stream (unit = node * n_mats + m)
struct GemvState { int st; uint32_t ph; int wb; uint32_t wph; int rb; uint32_t rph; };
struct EwState { int vs; uint32_t vph; int wb; uint32_t wph; int rb; uint32_t rph; };
struct EwIter { float a1, a2, sc1, sc2; int br1, br2, cur; };

__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// ---- loader warp: matrices through the TMA ring, dual vectors through the cp.async vector ring -----------------
struct LoaderState { int st; uint32_t ph; int vs; uint32_t vph; int skip; };
__device__ __noinline__ void loader_role(const PArgs &P, const Slice &R, LoaderState &L, int it) {
    const Pipe M = pipe_smem();
    const int nx = P.nx, nu = P.nu, nv = P.nv, ny = 2 * nx + nu, lane = threadIdx.x & 31;
    int st = L.st, vs = L.vs; uint32_t ph = L.ph, vph = L.vph;
    const int skip = L.skip;   // chunks of this iteration that were issued ahead, before the sweeps of the previous one
    long long cyc_empty = 0, cyc_vec = 0, cyc_go = 0;
    const bool pair_ok = (nu & 1) == 0;   // 8-byte copies: every vector starts on an 8-byte boundary when nu is even
    auto load_vec = [&](int node, int it) {
        const int prev = (it + 1) & 1;   // W[(it-1)&1] and Y[(it&1)^1]
        const size_t ox = (size_t)node * 2 * nx, op = (size_t)node * nu;
        const long long cv_ = clock64();
        mbar_wait(&M.vempty[vs], vph ^ 1);
        float *dst = M.vec + vs * kVecCount * kVStride;
        const float *s0 = P.pri_xi + ox, *s1 = P.Wxi[prev] + ox, *s2 = P.dual_xi + ox, *s3 = P.Yxi[prev] + ox;
        const float *s8 = P.diag + (size_t)node * ny;
        const float *s4 = P.pri_psi + op, *s5 = P.Wpsi[prev] + op, *s6 = P.dual_psi + op, *s7 = P.Ypsi[prev] + op;
        if (pair_ok) {
            for (int k = 2 * lane; k < 2 * nx; k += 64) {
                cp_async8(dst + k, s0 + k); cp_async8(dst + kVStride + k, s1 + k); cp_async8(dst + 2 * kVStride + k, s2 + k);
                cp_async8(dst + 3 * kVStride + k, s3 + k); cp_async8(dst + 8 * kVStride + k, s8 + k);
            }
            for (int k = 2 * lane; k < nu; k += 64) {
                cp_async8(dst + 4 * kVStride + k, s4 + k); cp_async8(dst + 5 * kVStride + k, s5 + k);
                cp_async8(dst + 6 * kVStride + k, s6 + k); cp_async8(dst + 7 * kVStride + k, s7 + k);
            }
        } else {
            for (int k = lane; k < 2 * nx; k += 32) {
                cp_async4(dst + k, s0 + k); cp_async4(dst + kVStride + k, s1 + k); cp_async4(dst + 2 * kVStride + k, s2 + k);
                cp_async4(dst + 3 * kVStride + k, s3 + k); cp_async4(dst + 8 * kVStride + k, s8 + k);
            }
            for (int k = lane; k < nu; k += 32) {
                cp_async4(dst + 4 * kVStride + k, s4 + k); cp_async4(dst + 5 * kVStride + k, s5 + k);
                cp_async4(dst + 6 * kVStride + k, s6 + k); cp_async4(dst + 7 * kVStride + k, s7 + k);
            }
        }
        cp_async_mbar_arrive(&M.vfull[vs]);
        if (++vs == kVecSlots) { vs = 0; vph ^= 1; }
        cyc_vec += clock64() - cv_;
    };
    // issue the chunks of units [u_begin, u_end) in order; the first `skip_n` are already in flight; stop after
    // `limit` issued chunks (prefetch).  Returns the number of chunks issued.
    auto stream_chunks = [&](int it, int skip_n, int limit, bool with_vec) {
        int issued = 0;
        for (int u = R.u_begin; u < R.u_end; u++) {
            const int node = u / P.n_mats, m = u - node * P.n_mats;
            if (with_vec && (u == R.u_begin || m == 0) && node + 2 <= R.node_last) load_vec(node + 2, it);
            const int ncols = (m & 1) == 0 ? 2 * nx : nu;
            const float *base = P.mat[m] + (size_t)node * nv * ncols;
            for (int c0 = 0; c0 < ncols; c0 += P.cols_per_chunk) {
                if (skip_n > 0) { skip_n--; continue; }
                if (issued >= limit) return issued;
                const int cc = min(P.cols_per_chunk, ncols - c0);
                const uintptr_t p0 = reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv);
                const uintptr_t p1 = p0 + (size_t)cc * nv * sizeof(float);
                const uintptr_t b0 = p0 & ~uintptr_t(15), b1 = (p1 + 15) & ~uintptr_t(15);
                const uint32_t bytes = (uint32_t)(b1 - b0);
                { const long long c_ = clock64(); mbar_wait(&M.empty[st], ph ^ 1); cyc_empty += clock64() - c_; }
                if (lane == 0) {
                    mbar_expect_tx(&M.full[st], bytes);
                    bulk_g2s(M.ring + st * kPStageStride, reinterpret_cast<const void *>(b0), bytes, &M.full[st]);
                }
                __syncwarp();
                if (++st == kPStages) { st = 0; ph ^= 1; }
                issued++;
            }
        }
        return issued;
    };
    const long long cg_ = clock64();
    if (R.node_first <= R.node_last) load_vec(R.node_first, it);
    if (R.node_first + 1 <= R.node_last) load_vec(R.node_first + 1, it);
    stream_chunks(it, skip, 0x7fffffff, true);
    // the matrices are constant: the next iteration's first ring-full is requested now and lands while the sweeps run
    L.skip = it + 1 < P.iters ? stream_chunks(it + 1, 0, kPStages, false) : 0;
    L.st = st; L.ph = ph; L.vs = vs; L.vph = vph;
    cyc_go += clock64() - cg_;
    if (blockIdx.x == 0 && lane == 0) { P.phase_ns[28] += cyc_empty; P.phase_ns[29] += cyc_vec; P.phase_ns[23] += cyc_go; }
}

// ---- GEMV warps: part[m][node] = (factor matrix m of the node) x (w segment), warp = column, lane = rows lane + 32 k.
// Rows >= nv of the last 32-row group read the neighbouring column (or the stage's slack): those lanes' sums are
// never used, so the inner loop carries no row predicate.
template <int NR>
__device__ __forceinline__ void gemv_role(const PArgs &P, const Slice &R, GemvState &G) {
    const Pipe M = pipe_smem();
    const int nx = P.nx, nu = P.nu, nv = P.nv, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int st = G.st, wb = G.wb, rb = G.rb; uint32_t ph = G.ph, wph = G.wph, rph = G.rph;
    int node_prev = -1;
    const bool dbg = blockIdx.x == 0 && threadIdx.x == 0;
    long long cyc_full = 0, cyc_w = 0, cyc_red = 0, cyc_cmp = 0;
    for (int u = R.u_begin; u < R.u_end; u++) {
        const int node = u / P.n_mats, m = u - node * P.n_mats;
        if (node != node_prev) {
            if (node_prev >= 0) {   // done with the previous node's w
                __syncwarp();
                if (lane == 0) mbar_arrive(&M.wempty[wb]);
                if (++wb == 2) { wb = 0; wph ^= 1; }
            }
            { const long long c_ = clock64(); mbar_wait(&M.wfull[wb], wph); cyc_w += clock64() - c_; }
            node_prev = node;
        }
        const bool xi_type = (m & 1) == 0;
        const int ncols = xi_type ? 2 * nx : nu;
        const float *wseg = M.wbuf + wb * kWStride + (xi_type ? 0 : kVStride);
        const float *base = P.mat[m] + (size_t)node * nv * ncols;
        float acc[NR];
#pragma unroll
        for (int k = 0; k < NR; k++) acc[k] = 0.f;
        for (int c0 = 0; c0 < ncols; c0 += P.cols_per_chunk) {
            const int cc = min(P.cols_per_chunk, ncols - c0);
            const int off = (int)((reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv) & 15) >> 2);
            const long long c1_ = clock64();
            mbar_wait(&M.full[st], ph);
            const long long c2_ = clock64();
            cyc_full += c2_ - c1_;
            const float *sb = M.ring + st * kPStageStride + off + lane;
            const float *wc = wseg + c0;
            int j = warp;
            for (; j + kGemvWarps < cc; j += 2 * kGemvWarps) {   // two columns per trip: all loads first
                const float w0 = wc[j], w1 = wc[j + kGemvWarps];
                const float *col0 = sb + j * nv, *col1 = col0 + kGemvWarps * nv;
                float a[NR], b[NR];
#pragma unroll
                for (int k = 0; k < NR; k++) { a[k] = col0[32 * k]; b[k] = col1[32 * k]; }
#pragma unroll
                for (int k = 0; k < NR; k++) { acc[k] = fmaf(a[k], w0, acc[k]); acc[k] = fmaf(b[k], w1, acc[k]); }
            }
            if (j < cc) {
                const float w0 = wc[j];
                const float *col0 = sb + j * nv;
                float a[NR];
#pragma unroll
                for (int k = 0; k < NR; k++) a[k] = col0[32 * k];
#pragma unroll
                for (int k = 0; k < NR; k++) acc[k] = fmaf(a[k], w0, acc[k]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&M.empty[st]);
            if (++st == kPStages) { st = 0; ph ^= 1; }
            cyc_cmp += clock64() - c2_;
        }
        // hand the per-warp partial sums to the element-wise warps
        { const long long c_ = clock64(); mbar_wait(&M.rempty[rb], rph ^ 1); cyc_red += clock64() - c_; }
        float *rd = M.red + (rb * kGemvWarps + warp) * kDimMax + lane;
#pragma unroll
        for (int k = 0; k < NR; k++) rd[32 * k] = acc[k];
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.rfull[rb]);
        if (++rb == 2) { rb = 0; rph ^= 1; }
    }
    if (node_prev >= 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.wempty[wb]);
        if (++wb == 2) { wb = 0; wph ^= 1; }
    }
    G.st = st; G.ph = ph; G.wb = wb; G.wph = wph; G.rb = rb; G.rph = rph;
    if (dbg) { P.phase_ns[24] += cyc_full; P.phase_ns[25] += cyc_w; P.phase_ns[26] += cyc_red; P.phase_ns[27] += cyc_cmp; }
}

// ---- element-wise warps: fused finalisation (previous iteration) + extrapolation (this one), one node ahead of the
// GEMV warps; then the cross-warp reduction of the GEMV partial sums.
__device__ __noinline__ void ew_role(const PArgs &P, const Slice &R, EwState &E, const EwIter &I, Cand &bx, Cand &bp) {
    const Pipe M = pipe_smem();
    const int nx = P.nx, nu = P.nu, nv = P.nv, ny = 2 * nx + nu, lane = threadIdx.x & 31;
    const int et = threadIdx.x - kGemvWarps * 32;   // thread index among the element-wise warps
    int vs = E.vs, wb = E.wb, rb = E.rb; uint32_t vph = E.vph, wph = E.wph, rph = E.rph;
    long long cyc_rf = 0, cyc_pro = 0, cyc_vf = 0;
    auto prologue = [&](int node) {
        const long long cp_ = clock64();
        const bool wr_xi = R.u_begin <= node * P.n_mats && node * P.n_mats < R.u_end;           // owner of (node, D)
        const bool wr_psi = R.u_begin <= node * P.n_mats + 1 && node * P.n_mats + 1 < R.u_end;   // owner of (node, F)
        mbar_wait(&M.vfull[vs], vph);
        mbar_wait(&M.wempty[wb], wph ^ 1);
        cyc_vf += clock64() - cp_;
        const float *vsl = M.vec + vs * kVecCount * kVStride;
        float *wdst = M.wbuf + wb * kWStride;
        for (int el = et; el < ny; el += kEwWarps * 32) {
            const bool xi_type = el < 2 * nx;
            const int t = xi_type ? el : el - 2 * nx;
            const float *v4 = vsl + (xi_type ? 0 : 4 * kVStride) + t;
            const float hx = v4[0], wp = v4[kVStride], yp = v4[3 * kVStride];
            float z = v4[2 * kVStride];
            if (xi_type && (I.br1 || I.br2)) {   // distance branch of the previous prox (:792-815; quirk SURVEY A.4-1)
                const float tt = hx + P.inv_step * wp;
                const float df = tt + -1.f * z;
                if (t < nx) { if (I.br1) z = z + I.sc1 * df; }
                else if (I.br2) { const float d2v = I.br1 ? (node == 0 ? 0.f : df + -1.f * z) : df; z = z + I.sc2 * d2v; }
            }
            const float res = hx + -1.f * z;            // computeFixedPointResidual (:839-850)
            const float yn = wp + P.step * res;         // dualUpdate (:854-864)
            float w = yn * I.a1;                        // dualExtrapolationStep (:548-552)
            w += I.a2 * yp;
            wdst[(xi_type ? 0 : kVStride) + t] = w;
            if (xi_type ? wr_xi : wr_psi) {
                const size_t k = xi_type ? (size_t)node * 2 * nx + t : (size_t)node * nu + t;
                if (xi_type) { P.Yxi[I.cur][k] = yn; P.Wxi[I.cur][k] = w; } else { P.Ypsi[I.cur][k] = yn; P.Wpsi[I.cur][k] = w; }
                const Cand cd{fabsf(res), res, (int)k};
                if (xi_type) cand_merge(bx, cd); else cand_merge(bp, cd);
            }
        }
        ewbar();
        if (wr_xi) {   // c = sysF' xi_w  (:651-658)
            const float *dg = vsl + 8 * kVStride;
            for (int t = et; t < nx; t += kEwWarps * 32) P.c[(size_t)node * nx + t] = dg[t] * wdst[t] + dg[nx + t] * wdst[nx + t];
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&M.vempty[vs]); mbar_arrive(&M.wfull[wb]); }
        if (++vs == kVecSlots) { vs = 0; vph ^= 1; }
        if (++wb == 2) { wb = 0; wph ^= 1; }
        cyc_pro += clock64() - cp_;
    };
    if (R.node_first <= R.node_last) prologue(R.node_first);
    for (int u = R.u_begin; u < R.u_end; u++) {
        const int node = u / P.n_mats, m = u - node * P.n_mats;
        if ((u == R.u_begin || m == 0) && node < R.node_last) prologue(node + 1);   // one node ahead of the GEMV warps
        { const long long c_ = clock64(); mbar_wait(&M.rfull[rb], rph); cyc_rf += clock64() - c_; }
        if (et < nv) {
            const float *rd = M.red + rb * kGemvWarps * kDimMax + et;
            float sum = rd[0];
#pragma unroll
            for (int w = 1; w < kGemvWarps; w++) sum += rd[w * kDimMax];
            P.part[m][(size_t)node * nv + et] = sum;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.rempty[rb]);
        if (++rb == 2) { rb = 0; rph ^= 1; }
    }
    E.vs = vs; E.vph = vph; E.wb = wb; E.wph = wph; E.rb = rb; E.rph = rph;
    if (blockIdx.x == 0 && et == 0) { P.phase_ns[30] += cyc_rf; P.phase_ns[31] += cyc_pro; P.phase_ns[22 + 0] += 0; }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPT, 1) k_apg_persistent(const __grid_constant__ PArgs P) {
    const int nv = P.nv;
    const Pipe M = pipe_smem();
    double *dsh = reinterpret_cast<double *>(smem_f(kOffDsh));      // 2 * 16 doubles
    Cand *csh = reinterpret_cast<Cand *>(smem_f(kOffCsh));          // 2 * 16 candidates
    float *sd = smem_f(kOffSd);                                     // d1, d2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kPStages; s++) { mbar_init(&M.full[s], 1); mbar_init(&M.empty[s], kGemvWarps); }
        for (int s = 0; s < kVecSlots; s++) { mbar_init(&M.vfull[s], 32); mbar_init(&M.vempty[s], kEwWarps); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&M.wfull[s], kEwWarps); mbar_init(&M.wempty[s], kGemvWarps);
            mbar_init(&M.rfull[s], kGemvWarps); mbar_init(&M.rempty[s], kEwWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // this CTA's slice of the stream: matrix units [u_begin, u_end) in node-major order
    const long long n_units = (long long)P.nodes * P.n_mats;
    Slice R;
    R.u_begin = (int)(n_units * blockIdx.x / gridDim.x);
    R.u_end = (int)(n_units * (blockIdx.x + 1) / gridDim.x);
    R.node_first = R.u_begin < R.u_end ? R.u_begin / P.n_mats : 0;
    R.node_last = R.u_begin < R.u_end ? (R.u_end - 1) / P.n_mats : -1;

    unsigned int bar_target = 0;
    GemvState GS{0, 0u, 0, 0u, 0, 0u};
    EwState ES{0, 0u, 0, 0u, 0, 0u};
    LoaderState LS{0, 0u, 0, 0u, 0};
    double s1 = 0, s2 = 0;
    if (blockIdx.x == 0 && tid == 0) g_t_prev = globaltimer();
    auto stamp = [&](int idx) { dstamp(P.phase_ns, idx); };
    const bool is_gemv = warp < kGemvWarps;
    const int nr = (nv + 31) >> 5;

    for (int it = 0; it < P.iters; it++) {
        const float lam = __ldg(P.lambda_tab + it);
        const int cur = it & 1;
        // ---- global distances of the previous iteration's prox (cublasSnrm2, :792, :810); zeros at it == 0
        {
            double p1 = 0, p2 = 0;
            for (int k = tid; k < (int)gridDim.x; k += kPC) { p1 += __ldcg(P.dist_part + 2 * k); p2 += __ldcg(P.dist_part + 2 * k + 1); }
            for (int o = 16; o > 0; o >>= 1) { p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o); }
            if (lane == 0) { dsh[warp] = p1; dsh[kPC / 32 + warp] = p2; }
            cbar();
            if (tid == 0) {
                double t1 = 0, t2 = 0;
                for (int w = 0; w < kPC / 32; w++) { t1 += dsh[w]; t2 += dsh[kPC / 32 + w]; }
                sd[0] = (float)sqrt(t1); sd[1] = (float)sqrt(t2);
            }
            cbar();
        }
        Cand bx{-1.f, 0.f, 0x7fffffff}, bp{-1.f, 0.f, 0x7fffffff};

        // ---- phase S: three-role pipeline.  loader -> [ring] -> GEMV warps -> [red] -> element-wise warps -> [wbuf] -> GEMV
        if (is_gemv) {
            if (nr == 4) gemv_role<4>(P, R, GS); else if (nr == 3) gemv_role<3>(P, R, GS);
            else if (nr == 2) gemv_role<2>(P, R, GS); else gemv_role<1>(P, R, GS);
        } else if (warp == kLoaderWarp) {
            loader_role(P, R, LS, it);
        } else {
            const float d1 = sd[0], d2 = sd[1];
            const float thr1 = P.inv_step * P.pen_x, thr2 = P.inv_step * P.pen_xs;
            EwIter I;
            I.a1 = 1.f + lam; I.a2 = -lam; I.cur = cur;
            I.br1 = d1 > thr1; I.br2 = d2 > thr2;
            I.sc1 = I.br1 ? 1.f - thr1 / d1 : 0.f; I.sc2 = I.br2 ? 1.f - thr2 / d2 : 0.f;
            ew_role(P, R, ES, I, bx, bp);
        }
        // infeasibility candidates of iteration it-1 (updatePrimalInfeasibity, :1480-1496)
        if (it > 0) {
            bx = cand_warp(bx); bp = cand_warp(bp);
            if (lane == 0) { csh[warp] = bx; csh[kPC / 32 + warp] = bp; }
            cbar();
            if (tid == 0) {
                Cand x = csh[0], p = csh[kPC / 32];
                for (int w = 1; w < kPC / 32; w++) { cand_merge(x, csh[w]); cand_merge(p, csh[kPC / 32 + w]); }
                float *o = P.pinf_part + 6 * (size_t)blockIdx.x;
                o[0] = x.a; o[1] = x.v; o[2] = __int_as_float(x.idx); o[3] = p.a; o[4] = p.v; o[5] = __int_as_float(p.idx);
            }
        }
        stamp(0);
        grid_sync(P.bar, bar_target);
        stamp(1);
        if (it > 0 && blockIdx.x == 0 && warp == 0) {
            Cand x{-1.f, 0.f, 0x7fffffff}, p{-1.f, 0.f, 0x7fffffff};
            for (int b = lane; b < (int)gridDim.x; b += 32) {
                const float *o = P.pinf_part + 6 * (size_t)b;
                Cand cx{__ldcg(o), __ldcg(o + 1), __float_as_int(__ldcg(o + 2))}, cp{__ldcg(o + 3), __ldcg(o + 4), __float_as_int(__ldcg(o + 5))};
                cand_merge(x, cx); cand_merge(p, cp);
            }
            x = cand_warp(x); p = cand_warp(p);
            if (lane == 0) P.pinf[it - 1] = fmaxf(x.v, p.v);    // max(maxValueXi, maxValuePsi) (:1495)
        }

        // ---- phase B: backward sweep
        stamp(2);
        for (int j = blockIdx.x; j < P.K && P.cs < P.N; j += gridDim.x) tail_backward(P, j);
        stamp(9);
        if (P.cs > 0 && P.cs < P.N) grid_sync(P.bar, bar_target);
        stamp(10);
        for (int s = P.cs - 1; s >= 0; s--) {
            const int first = __ldg(P.cum + s), last = __ldg(P.cum + s + 1);
            for (int i = first + blockIdx.x; i < last; i += gridDim.x) crown_backward(P, i);
            stamp(11);
            if (s > 0) grid_sync(P.bar, bar_target);
            stamp(12);
        }
        // ---- phase F: forward sweep + prox boxes
        const float *wxi = P.Wxi[cur], *wpsi = P.Wpsi[cur];
        for (int s = 0; s < P.cs; s++) {
            const int first = __ldg(P.cum + s), last = __ldg(P.cum + s + 1);
            const int branching = s > 0 && (last - first) > (first - __ldg(P.cum + s - 1));
            for (int i = first + blockIdx.x; i < last; i += gridDim.x) crown_forward(P, i, branching, wxi, wpsi, s1, s2);
            stamp(13);
            grid_sync(P.bar, bar_target);
            stamp(14);
        }
        for (int j = blockIdx.x; j < P.K && P.cs < P.N; j += gridDim.x) tail_forward(P, j, wxi, wpsi, s1, s2);
        stamp(20);
        {   // this CTA's share of the two squared distances
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            if (lane == 0) { dsh[warp] = s1; dsh[kPC / 32 + warp] = s2; }
            cbar();
            if (tid == 0) {
                double t1 = 0, t2 = 0;
                for (int w = 0; w < kPC / 32; w++) { t1 += dsh[w]; t2 += dsh[kPC / 32 + w]; }
                P.dist_part[2 * blockIdx.x] = t1; P.dist_part[2 * blockIdx.x + 1] = t2;
            }
            s1 = 0; s2 = 0;
        }
        stamp(21);
        grid_sync(P.bar, bar_target);
        stamp(22);
    }
    if (blockIdx.x == 0 && tid == 0) *P.iter_dev = P.iters - 1;   // k_finalize finishes iteration iters-1
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static size_t persist_smem_bytes(const Handle *) { return (size_t)kSmemFloats * 4 + 128; }

bool persistent_supported(const Handle *h) {
    const rn_dims &d = h->d;
    if (std::max(2 * d.nx, std::max(d.nu, d.nv)) > kDimMax) return false;
    if (h->chain_stage < d.N && d.N - h->chain_stage > kTMax) return false;
    if (kPStageFloats / d.nv < 1) return false;
    if (6 * kDimMax > kDimMax * kTP) return false;
    return persist_smem_bytes(h) <= 227 * 1024;
}

rn_status persistent_prepare(Handle *h) {
    if (h->persist_ready) return RN_OK;
    const rn_dims &d = h->d;
    const size_t n = d.nodes;
    RN_CHECK(dev_alloc(h, &h->part[0], n * d.nv)); RN_CHECK(dev_alloc(h, &h->part[1], n * d.nv));
    RN_CHECK(dev_alloc(h, &h->part[2], n * d.nv)); RN_CHECK(dev_alloc(h, &h->part[3], n * d.nv));
    RN_CHECK(dev_alloc(h, &h->LV, n * d.nu));
    RN_CHECK(dev_alloc(h, &h->grid_bar, 8));
    RN_CHECK(dev_alloc(h, &h->phase_ns, 32));
    const size_t smem = persist_smem_bytes(h);
    RN_CUDA(h, cudaFuncSetAttribute(k_apg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    RN_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_apg_persistent, kPT, smem));
    if (per_sm < 1) return fail(h, RN_ERR_INVALID, "persistent kernel does not fit on an SM (%zu B shared memory)", smem);
    int coop = 0;
    RN_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) return fail(h, RN_ERR_CUDA, "device does not support cooperative launch");
    h->persist_grid = std::min(h->sm_count, std::min(h->dist_slots, h->pinf_slots));
    h->persist_ready = true;
    return RN_OK;
}

// iterations 0 .. iters-1 up to (and excluding) the last finalisation; the caller runs k_finalize afterwards
rn_status persistent_launch(Handle *h, cudaStream_t st, int iters) {
    const rn_dims &d = h->d;
    PArgs P{};
    P.parent = h->t.parent; P.child_first = h->t.child_first; P.child_count = h->t.child_count; P.omega_idx = h->t.omega_idx;
    P.cum = h->cum_dev;
    P.N = d.N; P.cs = h->chain_stage; P.K = d.K; P.nodes = d.nodes;
    P.n_mats = h->factor_mode == RN_FACTORS_FULL ? 4 : 2;
    P.df_mode = h->factor_mode == RN_FACTORS_DF ? 1 : 0;
    P.iters = iters; P.nx = d.nx; P.nu = d.nu; P.nv = d.nv;
    P.cols_per_chunk = kPStageFloats / d.nv;
    P.mat[0] = h->D; P.mat[1] = h->F; P.mat[2] = h->Phi; P.mat[3] = h->Psi;
    P.Omega = h->Omega; P.Theta = h->Theta; P.G = h->G; P.L = h->L; P.B = h->B; P.diag = h->diag;
    P.beta = h->beta; P.uhat = h->uhat; P.e = h->e; P.xcur = h->xcur; P.uprev = h->uprev; P.uhat_prev = h->uhat_prev;
    P.sxmin = h->sxmin; P.sxmax = h->sxmax; P.sxs = h->sxs; P.sumin = h->sumin; P.sumax = h->sumax;
    P.Yxi[0] = h->yA_xi; P.Yxi[1] = h->yB_xi; P.Ypsi[0] = h->yA_psi; P.Ypsi[1] = h->yB_psi;
    P.Wxi[0] = h->wA_xi; P.Wxi[1] = h->wB_xi; P.Wpsi[0] = h->wA_psi; P.Wpsi[1] = h->wB_psi;
    P.pri_xi = h->pri_xi; P.pri_psi = h->pri_psi; P.dual_xi = h->dual_xi; P.dual_psi = h->dual_psi;
    for (int k = 0; k < 4; k++) P.part[k] = h->part[k];
    P.c = h->c; P.q = h->q; P.r = h->r; P.sigma = h->sigma; P.V = h->V; P.U = h->U; P.X = h->X; P.LV = h->LV;
    P.dist_part = h->dist_part; P.pinf = h->pinf; P.pinf_part = h->pinf_part;
    P.lambda_tab = h->lambda_tab; P.bar = h->grid_bar; P.iter_dev = h->iter_dev; P.phase_ns = h->phase_ns;
    P.step = h->step; P.inv_step = 1 / h->step; P.pen_x = h->pen_x; P.pen_xs = h->pen_xs;
    RN_CUDA(h, cudaMemsetAsync(h->grid_bar, 0, sizeof(unsigned int), st));
    RN_CUDA(h, cudaMemsetAsync(h->dist_part, 0, 2 * sizeof(double) * h->persist_grid, st));
    void *args[] = {&P};
    RN_CUDA(h, cudaLaunchCooperativeKernel((const void *)k_apg_persistent, dim3(h->persist_grid), dim3(kPT), args,
                                           persist_smem_bytes(h), st));
    return RN_OK;
}

}  // namespace rn
