// rn_persist.cu -- the whole APG loop of SmpcController::algorithmApg as ONE persistent cooperative kernel.
//
// Reference hot loop: /root/reference/src/SmpcController.cu:1500-1525 (about 430 launches per iteration).
// Here every iteration runs inside one resident grid (one CTA per SM, 16 consumer warps + 1 loader warp) with
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

constexpr int kPC = 512;                        // consumer threads
constexpr int kPT = kPC + 32;                   // + loader warp
constexpr int kPStages = 6;
constexpr int kPStageFloats = 4096;
constexpr int kPStageStride = kPStageFloats + 32;
constexpr int kVecSlots = 2;
constexpr int kVecCount = 5;                    // Hx, w_prev, z, y_prev, diag
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
    unsigned long long *phase_ns;               // [4] stream, backward, forward-crown, forward-chains (CTA 0's clock)
    float step, inv_step, pen_x, pen_xs;
};

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kPC) : "memory"); }

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
#pragma unroll 8
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

// acc[c] += sum_k M[row + k*ldm] * Xt[k*kTP + col0 + c]   (M global and constant, Xt shared, transposed right-hand side)
__device__ __forceinline__ void chain_gemm(const float *__restrict__ M, int ldm, int row, int kdim, const float *Xt, int col0,
                                           float (&acc)[kCG]) {
    const float *mp = M + row;
    const float *xp = Xt + col0;
#pragma unroll 8
    for (int k = 0; k < kdim; k++) {
        const float mv = __ldg(mp + (size_t)k * ldm);
        const float2 x0 = *reinterpret_cast<const float2 *>(xp + k * kTP);
        const float2 x1 = *reinterpret_cast<const float2 *>(xp + k * kTP + 2);
        const float2 x2 = *reinterpret_cast<const float2 *>(xp + k * kTP + 4);
        acc[0] = fmaf(mv, x0.x, acc[0]); acc[1] = fmaf(mv, x0.y, acc[1]);
        acc[2] = fmaf(mv, x1.x, acc[2]); acc[3] = fmaf(mv, x1.y, acc[3]);
        acc[4] = fmaf(mv, x2.x, acc[4]); acc[5] = fmaf(mv, x2.y, acc[5]);
    }
}

struct TailSmem {
    float *A0, *A1, *A2, *A3;   // four [kDimMax][kTP] arrays
    float *scr;                 // kPC floats
};

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
// chains (stages cs .. N-1 of scenario j): backward
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tail_backward(const PArgs &P, int j, const TailSmem &S) {
    const int nx = P.nx, nv = P.nv, nu = P.nu, t = threadIdx.x, T = P.N - P.cs;
    float *Qb = S.A0, *Y3 = S.A1, *Sg = S.A2, *Vt = S.A3;
    const int row = t & (kDimMax - 1), cg = t >> 7, col0 = cg * kCG;
    // 0. operands of the two scans, staged transposed ([element][stage]) by all threads: c -> Qb, beta -> Sg, D xi -> Vt
    for (int idx = t; idx < T * nx; idx += kPC) {
        const int s = idx / nx, e = idx - s * nx;
        Qb[e * kTP + s] = __ldcg(P.c + (size_t)(__ldg(P.cum + P.cs + s) + j) * nx + e);
    }
    for (int idx = t; idx < T * nv; idx += kPC) {
        const int s = idx / nv, e = idx - s * nv;
        const size_t k = (size_t)(__ldg(P.cum + P.cs + s) + j) * nv + e;
        Sg[e * kTP + s] = __ldg(P.beta + k);
        Vt[e * kTP + s] = __ldcg(P.part[0] + k);
    }
    cbar();
    // 1. q-scan (in place): q_bar = q of the child (0 at the leaf), q = sysF' xi_w + q_bar  (:651-658)
    if (t < nx) {
        float qrun = 0.f;
        for (int s = T - 1; s >= 0; s--) { const float cv = Qb[t * kTP + s]; Qb[t * kTP + s] = qrun; qrun = cv + qrun; }
        P.q[(size_t)(__ldg(P.cum + P.cs) + j) * nx + t] = qrun;   // head of the chain, for the crown
    }
    // D xi + ... second operand of the r-scan, fetched while the GEMM runs
    float af[kTMax];
    if (t < nv) {
#pragma unroll
        for (int s = 0; s < kTMax; s++)
            if (s < T) af[s] = __ldcg(P.part[1] + (size_t)(__ldg(P.cum + P.cs + s) + j) * nv + t);
    }
    cbar();
    // 2. Y3 = G Qb
    if (row < nv) {
        float acc[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        chain_gemm(P.G, nv, row, nx, Qb, col0, acc);
#pragma unroll
        for (int c = 0; c < kCG; c++) Y3[row * kTP + col0 + c] = acc[c];
    }
    cbar();
    // 3. r-scan (in place): sigma = beta + r_child (:599); r = ((sigma + D xi) + F psi) + G q_bar (:631-646)
    if (t < nv) {
        float rrun = 0.f;
#pragma unroll
        for (int s = kTMax - 1; s >= 0; s--)
            if (s < T) {
                const float sg = Sg[t * kTP + s] + rrun;
                rrun = ((sg + Vt[t * kTP + s]) + af[s]) + Y3[t * kTP + s];
                Sg[t * kTP + s] = P.df_mode ? rrun : sg;
                P.sigma[(size_t)(__ldg(P.cum + P.cs + s) + j) * nv + t] = sg;
            }
        P.r[(size_t)(__ldg(P.cum + P.cs) + j) * nv + t] = rrun;
    }
    cbar();
    // 4. v = ((-1/2 Omega sigma + Theta q_bar) + Psi psi) + Phi xi  (:604-627)   [df: v = -1/2 Omega r]
    if (row < nv) {
        const int oi = __ldg(P.omega_idx + __ldg(P.cum + P.cs) + j);   // one Omega/Theta per chain (Engine.cu:210-221)
        float a1[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a2[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        chain_gemm(P.Omega + (size_t)oi * nv * nv, nv, row, nv, Sg, col0, a1);
        if (!P.df_mode) chain_gemm(P.Theta + (size_t)oi * nv * nx, nv, row, nx, Qb, col0, a2);
#pragma unroll
        for (int c = 0; c < kCG; c++) {
            const int s = col0 + c;
            if (s < T) {
                const size_t k = (size_t)(__ldg(P.cum + P.cs + s) + j) * nv + row;
                float v;
                if (P.df_mode) v = -0.5f * a1[c];
                else v = ((-0.5f * a1[c] + a2[c]) + __ldcg(P.part[3] + k)) + __ldcg(P.part[2] + k);
                P.V[k] = v;
                Vt[row * kTP + s] = v;
            }
        }
    }
    cbar();
    // 5. LV = L V   (:701, :727 -- the reference does this GEMM in the forward sweep)
    if (row < nu) {
        float acc[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        chain_gemm(P.L, nu, row, nv, Vt, col0, acc);
#pragma unroll
        for (int c = 0; c < kCG; c++) {
            const int s = col0 + c;
            if (s < T) P.LV[(size_t)(__ldg(P.cum + P.cs + s) + j) * nu + row] = acc[c];
        }
    }
    cbar();
}

// chains: forward.  wxi/wpsi = this iteration's accelerated duals.
__device__ __forceinline__ void tail_forward(const PArgs &P, int j, const TailSmem &S, const float *wxi, const float *wpsi,
                                             double &s1, double &s2) {
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu, t = threadIdx.x, T = P.N - P.cs;
    float *Ut = S.A0, *Xt = S.A1;
    const int row = t & (kDimMax - 1), cg = t >> 7, col0 = cg * kCG;
    const int head = __ldg(P.cum + P.cs) + j;
    const int par0 = __ldg(P.parent + head);
    // the chain's first stage is a branching stage of the reference's forward loop when it has more nodes than its
    // parent stage (:699-719): same sums, different association
    const bool head_br = P.cs > 0 && (__ldg(P.cum + P.cs + 1) - __ldg(P.cum + P.cs)) > (__ldg(P.cum + P.cs) - __ldg(P.cum + P.cs - 1));
    float *Uh = S.A2, *Lv = S.A3;
    // 0. operands of the u-scan staged transposed by all threads
    for (int idx = t; idx < T * nu; idx += kPC) {
        const int s = idx / nu, e = idx - s * nu;
        const size_t k = (size_t)(__ldg(P.cum + P.cs + s) + j) * nu + e;
        Uh[e * kTP + s] = __ldg(P.uhat + k);
        Lv[e * kTP + s] = __ldcg(P.LV + k);
    }
    // operands of the x-scan, fetched early
    float ev[kTMax];
    float xrun = 0.f;
    if (t >= kDimMax && t < kDimMax + nx) {
        const int e = t - kDimMax;
#pragma unroll
        for (int s = 0; s < kTMax; s++)
            if (s < T) ev[s] = __ldg(P.e + (size_t)(__ldg(P.cum + P.cs + s) + j) * nx + e);
        xrun = par0 < 0 ? __ldg(P.xcur + e) : __ldcg(P.X + (size_t)par0 * nx + e);
    }
    cbar();
    // 6. u-scan: u = ((uhat + u_par) - uhat_par) + L v   (:722-728; below its first stage a chain never branches)
    if (t < nu) {
        float up = par0 < 0 ? __ldg(P.uprev + t) : __ldcg(P.U + (size_t)par0 * nu + t);
        float uhp = par0 < 0 ? __ldg(P.uhat_prev + t) : __ldg(P.uhat + (size_t)par0 * nu + t);
        for (int s = 0; s < T; s++) {
            const float uh = Uh[t * kTP + s], lv = Lv[t * kTP + s];
            const float u = (s == 0 && head_br) ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;
            P.U[(size_t)(__ldg(P.cum + P.cs + s) + j) * nu + t] = u;
            Ut[t * kTP + s] = u;
            up = u; uhp = uh;
        }
    }
    cbar();
    // 7. BU = B U   (:715, :736)
    float bu[kCG] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (row < nx) {
        chain_gemm(P.B, nx, row, nu, Ut, col0, bu);
#pragma unroll
        for (int c = 0; c < kCG; c++) Xt[row * kTP + col0 + c] = bu[c];
    }
    cbar();
    // 8. x-scan: x = (x_par + e) + B u   (:730-737)
    if (t >= kDimMax && t < kDimMax + nx) {
        const int e = t - kDimMax;
#pragma unroll
        for (int s = 0; s < kTMax; s++)
            if (s < T) {
                const float x = (s == 0 && head_br) ? xrun + (ev[s] + Xt[e * kTP + s]) : (xrun + ev[s]) + Xt[e * kTP + s];
                P.X[(size_t)(__ldg(P.cum + P.cs + s) + j) * nx + e] = x;
                Xt[e * kTP + s] = x;
                xrun = x;
            }
    }
    cbar();
    // 9. Hx, Hu, box projections, distance partials for every node of the chain
    for (int idx = t; idx < T * ny; idx += kPC) {
        const int s = idx / ny, el = idx - s * ny;
        const int i = __ldg(P.cum + P.cs + s) + j;
        const float v = el < 2 * nx ? Xt[(el < nx ? el : el - nx) * kTP + s] : Ut[(el - 2 * nx) * kTP + s];
        prox_element(P, i, el, v, wxi, wpsi, s1, s2);
    }
    cbar();
}

// ---------------------------------------------------------------------------------------------------------------
// crown (stages above the chains): one node per CTA and per call
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void crown_backward(const PArgs &P, int i, const TailSmem &S) {
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

__device__ __forceinline__ void crown_forward(const PArgs &P, int i, int branching, const TailSmem &S, const float *wxi,
                                              const float *wpsi, double &s1, double &s2) {
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
__global__ void __launch_bounds__(kPT, 1) k_apg_persistent(const __grid_constant__ PArgs P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nx = P.nx, nu = P.nu, nv = P.nv, ny = 2 * nx + nu;
    const int vstride = (max(2 * nx, nu) + 31) & ~31;
    float *ring = reinterpret_cast<float *>(smem_raw);
    float *vec = ring + kPStages * kPStageStride;
    float *wbuf = vec + kVecSlots * kVecCount * vstride;
    float *red = wbuf + vstride;
    float *tail = red + kPC;
    TailSmem TS;
    TS.A0 = tail; TS.A1 = TS.A0 + kDimMax * kTP; TS.A2 = TS.A1 + kDimMax * kTP; TS.A3 = TS.A2 + kDimMax * kTP;
    TS.scr = TS.A3 + kDimMax * kTP;
    double *dsh = reinterpret_cast<double *>(TS.scr + kPC);          // 2 * 16 doubles
    Cand *csh = reinterpret_cast<Cand *>(dsh + 2 * (kPC / 32));     // 2 * 16 candidates
    float *sd = reinterpret_cast<float *>(csh + 2 * (kPC / 32));    // d1, d2
    uint64_t *full = reinterpret_cast<uint64_t *>(sd + 4);
    uint64_t *empty = full + kPStages;
    uint64_t *vfull = empty + kPStages;
    uint64_t *vempty = vfull + kVecSlots;
    uint64_t *go = vempty + kVecSlots;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = P.nodes * P.n_mats;
    if (tid == 0) {
        for (int s = 0; s < kPStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kPC / 32); }
        for (int s = 0; s < kVecSlots; s++) { mbar_init(&vfull[s], 32); mbar_init(&vempty[s], kPC / 32); }
        mbar_init(go, 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kPC / 32) {
        // =================================== loader warp ===================================
        int st = 0; uint32_t ph = 0; int vs = 0; uint32_t vph = 0;
        bool prefetched = false;
        auto load_vec = [&](int u, int it) {
            const int node = u / P.n_mats, m = u - node * P.n_mats;
            const bool xi_type = (m & 1) == 0;
            const int len = xi_type ? 2 * nx : nu;
            const size_t off = xi_type ? (size_t)node * 2 * nx : (size_t)node * nu;
            const int prev = (it + 1) & 1;   // W[(it-1)&1] and Y[(it&1)^1]
            const float *src0 = (xi_type ? P.pri_xi : P.pri_psi) + off;
            const float *src1 = (xi_type ? P.Wxi[prev] : P.Wpsi[prev]) + off;
            const float *src2 = (xi_type ? P.dual_xi : P.dual_psi) + off;
            const float *src3 = (xi_type ? P.Yxi[prev] : P.Ypsi[prev]) + off;
            const float *src4 = P.diag + (size_t)node * ny;
            mbar_wait(&vempty[vs], vph ^ 1);
            float *dst = vec + vs * kVecCount * vstride;
            for (int k = lane; k < len; k += 32) {
                cp_async4(dst + k, src0 + k);
                cp_async4(dst + vstride + k, src1 + k);
                cp_async4(dst + 2 * vstride + k, src2 + k);
                cp_async4(dst + 3 * vstride + k, src3 + k);
                if (m == 0) cp_async4(dst + 4 * vstride + k, src4 + k);
            }
            cp_async_mbar_arrive(&vfull[vs]);
            if (++vs == kVecSlots) { vs = 0; vph ^= 1; }
        };
        auto load_mat = [&](int u) {
            const int node = u / P.n_mats, m = u - node * P.n_mats;
            const int ncols = (m & 1) == 0 ? 2 * nx : nu;
            const float *base = P.mat[m] + (size_t)node * nv * ncols;
            for (int c0 = 0; c0 < ncols; c0 += P.cols_per_chunk) {
                const int cc = min(P.cols_per_chunk, ncols - c0);
                const uintptr_t p0 = reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv);
                const uintptr_t p1 = p0 + (size_t)cc * nv * sizeof(float);
                const uintptr_t b0 = p0 & ~uintptr_t(15), b1 = (p1 + 15) & ~uintptr_t(15);
                const uint32_t bytes = (uint32_t)(b1 - b0);
                mbar_wait(&empty[st], ph ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&full[st], bytes);
                    bulk_g2s(ring + st * kPStageStride, reinterpret_cast<const void *>(b0), bytes, &full[st]);
                }
                __syncwarp();
                if (++st == kPStages) { st = 0; ph ^= 1; }
            }
        };
        for (int it = 0; it < P.iters; it++) {
            mbar_wait(go, it & 1);   // the vectors of iteration `it` exist only after the barrier that ends it-1
            const int u0 = blockIdx.x;
            if (u0 < n_units) load_vec(u0, it);
            int k = 0;
            for (int u = u0; u < n_units; u += gridDim.x, k++) {
                if (u + (int)gridDim.x < n_units) load_vec(u + gridDim.x, it);
                if (!(k == 0 && prefetched)) load_mat(u);
            }
            prefetched = false;
            if (it + 1 < P.iters && u0 < n_units) { load_mat(u0); prefetched = true; }   // runs under the sweeps
        }
        return;
    }

    // =================================== consumers ===================================
    int slots = (nv + 31) & ~31;
    const int G = kPC / slots;                       // column groups of the stream GEMV
    const int g = tid / slots, r0 = tid - g * slots;
    const bool active = g < G && r0 < nv;
    unsigned int bar_target = 0;
    int st = 0; uint32_t ph = 0; int vs = 0; uint32_t vph = 0;
    double s1 = 0, s2 = 0;
    unsigned long long t_prev = 0;
    const bool clock_cta = blockIdx.x == 0 && tid == 0 && P.phase_ns != nullptr;
    if (clock_cta) t_prev = globaltimer();
    auto stamp = [&](int phase) {
        if (clock_cta) { const unsigned long long now = globaltimer(); P.phase_ns[phase] += now - t_prev; t_prev = now; }
    };

    for (int it = 0; it < P.iters; it++) {
        const float lam = __ldg(P.lambda_tab + it);
        const float a1 = 1.f + lam, a2 = -lam;
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
                mbar_arrive(go);
            }
            cbar();
        }
        const float d1 = sd[0], d2 = sd[1];
        const float thr1 = P.inv_step * P.pen_x, thr2 = P.inv_step * P.pen_xs;
        const bool br1 = d1 > thr1, br2 = d2 > thr2;
        const float sc1 = br1 ? 1.f - thr1 / d1 : 0.f, sc2 = br2 ? 1.f - thr2 / d2 : 0.f;
        Cand bx{-1.f, 0.f, 0x7fffffff}, bp{-1.f, 0.f, 0x7fffffff};

        // ---- phase S: fused element-wise pass + factor stream
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int node = u / P.n_mats, m = u - node * P.n_mats;
            const bool xi_type = (m & 1) == 0, writer = m < 2;
            const int ncols = xi_type ? 2 * nx : nu;
            mbar_wait(&vfull[vs], vph);
            const float *vsl = vec + vs * kVecCount * vstride;
            for (int t = tid; t < ncols; t += kPC) {
                const float hx = vsl[t], wp = vsl[vstride + t], yp = vsl[3 * vstride + t];
                float z = vsl[2 * vstride + t];
                if (xi_type && (br1 || br2)) {   // distance branch of the previous prox (:792-815; quirk SURVEY A.4-1)
                    const float tt = hx + P.inv_step * wp;
                    const float df = tt + -1.f * z;
                    if (t < nx) { if (br1) z = z + sc1 * df; }
                    else if (br2) { const float d2v = br1 ? (node == 0 ? 0.f : df + -1.f * z) : df; z = z + sc2 * d2v; }
                }
                const float res = hx + -1.f * z;            // computeFixedPointResidual (:839-850)
                const float yn = wp + P.step * res;         // dualUpdate (:854-864)
                float w = yn * a1;                          // dualExtrapolationStep (:548-552)
                w += a2 * yp;
                wbuf[t] = w;
                if (writer) {
                    const size_t k = xi_type ? (size_t)node * 2 * nx + t : (size_t)node * nu + t;
                    if (xi_type) { P.Yxi[cur][k] = yn; P.Wxi[cur][k] = w; } else { P.Ypsi[cur][k] = yn; P.Wpsi[cur][k] = w; }
                    const Cand cd{fabsf(res), res, (int)k};
                    if (xi_type) cand_merge(bx, cd); else cand_merge(bp, cd);
                }
            }
            cbar();
            if (m == 0) {   // c = sysF' xi_w  (:651-658)
                const float *dg = vsl + 4 * vstride;
                for (int t = tid; t < nx; t += kPC) P.c[(size_t)node * nx + t] = dg[t] * wbuf[t] + dg[nx + t] * wbuf[nx + t];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&vempty[vs]);
            if (++vs == kVecSlots) { vs = 0; vph ^= 1; }

            float acc = 0.f;
            const float *base = P.mat[m] + (size_t)node * nv * ncols;
            for (int c0 = 0; c0 < ncols; c0 += P.cols_per_chunk) {
                const int cc = min(P.cols_per_chunk, ncols - c0);
                const int off = (int)((reinterpret_cast<uintptr_t>(base + (size_t)c0 * nv) & 15) >> 2);
                mbar_wait(&full[st], ph);
                const float *sb = ring + st * kPStageStride + off + r0;
                if (active) {
#pragma unroll 4
                    for (int j = g; j < cc; j += G) acc = fmaf(sb[j * nv], wbuf[c0 + j], acc);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                if (++st == kPStages) { st = 0; ph ^= 1; }
            }
            if (active) red[g * nv + r0] = acc;
            cbar();
            if (tid < nv) {
                float s = red[tid];
                for (int gg = 1; gg < G; gg++) s += red[gg * nv + tid];
                P.part[m][(size_t)node * nv + tid] = s;
            }
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
        grid_sync(P.bar, bar_target);
        stamp(0);
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
        for (int j = blockIdx.x; j < P.K && P.cs < P.N; j += gridDim.x) tail_backward(P, j, TS);
        if (P.cs > 0 && P.cs < P.N) grid_sync(P.bar, bar_target);
        for (int s = P.cs - 1; s >= 0; s--) {
            const int first = __ldg(P.cum + s), last = __ldg(P.cum + s + 1);
            for (int i = first + blockIdx.x; i < last; i += gridDim.x) crown_backward(P, i, TS);
            if (s > 0) grid_sync(P.bar, bar_target);
        }
        stamp(1);
        // ---- phase F: forward sweep + prox boxes
        const float *wxi = P.Wxi[cur], *wpsi = P.Wpsi[cur];
        for (int s = 0; s < P.cs; s++) {
            const int first = __ldg(P.cum + s), last = __ldg(P.cum + s + 1);
            const int branching = s > 0 && (last - first) > (first - __ldg(P.cum + s - 1));
            for (int i = first + blockIdx.x; i < last; i += gridDim.x) crown_forward(P, i, branching, TS, wxi, wpsi, s1, s2);
            grid_sync(P.bar, bar_target);
        }
        stamp(2);
        for (int j = blockIdx.x; j < P.K && P.cs < P.N; j += gridDim.x) tail_forward(P, j, TS, wxi, wpsi, s1, s2);
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
        grid_sync(P.bar, bar_target);
        stamp(3);
    }
    if (blockIdx.x == 0 && tid == 0) *P.iter_dev = P.iters - 1;   // k_finalize finishes iteration iters-1
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static size_t persist_smem_bytes(const Handle *h) {
    const int vstride = (std::max(2 * h->d.nx, h->d.nu) + 31) & ~31;
    size_t f = (size_t)kPStages * kPStageStride + (size_t)kVecSlots * kVecCount * vstride + vstride + kPC +
               4 * (size_t)kDimMax * kTP + kPC;
    size_t b = f * 4 + 2 * (kPC / 32) * sizeof(double) + 2 * (kPC / 32) * sizeof(Cand) + 4 * sizeof(float) +
               (2 * kPStages + 2 * kVecSlots + 1) * sizeof(uint64_t);
    return b + 128;
}

bool persistent_supported(const Handle *h) {
    const rn_dims &d = h->d;
    if (std::max(d.nx, std::max(d.nu, d.nv)) > kDimMax) return false;
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
    RN_CHECK(dev_alloc(h, &h->phase_ns, 8));
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
