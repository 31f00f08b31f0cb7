// rn_persist.cu -- the whole APG loop of SmpcController::algorithmApg as ONE persistent cooperative kernel.
//
// Reference hot loop: /root/reference/src/SmpcController.cu:1500-1525 (about 430 launches per iteration).
// Here every iteration runs inside one resident grid (one CTA per SM, 16 warps) with four
// software grid barriers per iteration ("persistent kernel"):
//
//   phase S  factor stream.  Work unit = (node, matrix in {D, F, Phi, Psi}).  The loader warp keeps a 6-deep ring of
//            16 KB stages full with 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP) of the node's packed Engine
//            factor matrix and fetches the unit's dual vectors with cp.async into a 3-deep vector ring, one unit
//            ahead.  The consumer prologue is the fused element-wise pass: finalisation of the PREVIOUS iteration's prox
//            (distance branch), fixed-point residual, dual update y+ = w + step*res, infeasibility log, and the Nesterov
//            extrapolation w = (1+l) y+ - l y of THIS iteration (:535-557, :792-864, :1480-1496) -- the duals are read
//            once and written once per iteration.  Then partial products D xi_w, F psi_w, Phi xi_w, Psi psi_w.
//   sweeps   (:593-747).  The shared matrices G = Bbar', [OmegaBar | ThetaBar], L and B (every Omega_i / Theta_i is
//            OmegaBar / p_i, ThetaBar / p_i: Engine.cu:707-747) are pulled into shared memory once per iteration by
//            four bulk TMA copies that overlay the (then idle) stream ring, so every sweep product is a GEMM
//            [matrix in smem] x [24 columns in smem] across nodes.
//     phase B  chains: below the last branching stage every scenario is an independent chain owned by one CTA; the
//            stage recursion becomes scans (q = c + q_child; sigma = beta + r_child, r = sigma + D xi + F psi +
//            G q_child) around GEMMs across the chain's stages -- no barrier per stage.
//     phase C  crown (stages above the chains): the child->parent sums of solveSumChildren (Utilities.cu:168-201) are
//            unrolled into sums over each node's descendants (contiguous id ranges per stage), so all crown nodes are
//            independent and the crown costs ONE barrier instead of one per stage.
//     phase F  forward sweep (:675-747): u and x are path sums from the root; crown nodes and chains run in the same
//            phase (a chain recomputes its parent's u, x from the crown's L v).  Epilogue: Hx = sysF x, Hu = sysG u,
//            t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial sums of the two global
//            distances of proximalFunG (:792, :810).
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
constexpr int kTP = 24;                         // columns of a sweep tile (a chain's stages, or 24 crown nodes)
constexpr int kMaxCs = 8;                       // deepest crown supported (stages above the chains)
constexpr int kDimMax = 128;                    // max(2nx, nu, nv) supported by this kernel

struct PArgs {
    const int *parent, *child_first, *child_count, *omega_idx, *cum, *stages;
    const int *crown_rng;                       // [n_crown][kMaxCs + 1][2]: descendant id range of a crown node per stage
    int N, cs, K, nodes, n_crown, n_mats, df_mode, iters, nx, nu, nv, cols_per_chunk, clock_cta;
    const float *mat[4];                        // D, F, Phi, Psi (packed per node, Engine.cu:201-207)
    const float *pack;                          // G | OmegaBar ThetaBar | L | B, each padded to 16 B (sweeps)
    const float *diag, *prob;
    const float *beta, *uhat, *e, *xcur, *uprev, *uhat_prev, *sxmin, *sxmax, *sxs, *sumin, *sumax;
    float *Yxi[2], *Ypsi[2], *Wxi[2], *Wpsi[2];
    float *pri_xi, *pri_psi, *dual_xi, *dual_psi;
    float *part[4];                             // D xi_w, F psi_w, Phi xi_w, Psi psi_w   [nodes*nv] each
    float *c, *qh, *rh, *sigma, *V, *U, *X, *LV; // qh, rh: q and r of the chain heads  [K*nx], [K*nv]
    double *dist_part;                          // [2*grid]
    float *pinf, *pinf_part;                    // [iters], [grid*6]
    const float *lambda_tab;
    unsigned int *bar;
    int *iter_dev;
    unsigned long long *phase_ns;               // [32] fine-grained phase clock of one CTA (see cabi.PHASE_NAMES)
    float step, inv_step, pen_x, pen_xs;
    // sweep shared-memory layout (float offsets from the dynamic shared-memory base) and the pack's pieces
    int oG, oOT, oL, oB, oX1, oY, oV, oScr2;
    unsigned int bG, bOT, bL, bB;               // bytes of the four bulk copies
    int pG, pOT, pL, pB;                        // float offsets inside the pack
};

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kPC) : "memory"); }
__device__ __forceinline__ void ewbar() { asm volatile("bar.sync 2, %0;" ::"n"(kEwWarps * 32) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
// fine-grained clock of one CTA (thread 0): accumulates the time since the previous stamp into phase_ns[idx]
__device__ unsigned long long g_t_prev;
__device__ __forceinline__ void dstamp(const PArgs &P, int idx) {
    if (blockIdx.x == P.clock_cta && threadIdx.x == 0) {
        const unsigned long long now = globaltimer();
        P.phase_ns[idx] += now - g_t_prev;
        g_t_prev = now;
    }
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid barrier over all CTAs (the grid is co-resident: cooperative launch).  Same protocol as
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

// shared-memory layout (float offsets from the dynamic shared-memory base).  The small persistent area comes first;
// the phase-S region and the sweep region (runtime offsets in PArgs, they depend on nx/nu/nv) overlay each other
// behind it.  Phase-S offsets are compile-time constants so that every access there is an LDS/STS with an immediate
// offset even inside the non-inlined role functions.
constexpr int kOffScr = 0;                                              // kPC floats
constexpr int kOffDsh = kOffScr + kPC;                                  // 2 x 16 doubles
constexpr int kOffCsh = kOffDsh + 2 * 2 * (kPC / 32);                   // 2 x 16 candidates (3 words each)
constexpr int kOffSd = kOffCsh + 3 * 2 * (kPC / 32);                    // d1, d2
constexpr int kOffMisc = kOffSd + 4;                                    // 64 ints: column -> node map etc.
constexpr int kOffBar = kOffMisc + 64;                                  // mbarriers
constexpr int kNumBars = 2 * kPStages + 2 * kVecSlots + 8 + 4;
constexpr int kOffRing = (kOffBar + 2 * kNumBars + 31) & ~31;           // 128-byte aligned: TMA destination
constexpr int kOffVec = kOffRing + kPStages * kPStageStride;            // kVecSlots x kVecCount x kVStride
constexpr int kOffW = kOffVec + kVecSlots * kVecCount * kVStride;       // 2 x kWStride  (xi part | psi part)
constexpr int kOffRed = kOffW + 2 * kWStride;                           // 2 x kGemvWarps x kDimMax
constexpr int kStreamEnd = kOffRed + 2 * kGemvWarps * kDimMax;
constexpr int kOffSweep = kOffRing;                                     // the sweep region starts where the ring starts
static_assert(kOffDsh % 2 == 0 && kOffBar % 2 == 0, "8-byte alignment of the double / mbarrier areas");

__device__ __forceinline__ float *smem_f(int off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    return reinterpret_cast<float *>(smem_raw) + off;
}

// ---------------------------------------------------------------------------------------------------------------
// sweeps.  A "tile" is up to kTP columns that go through the same shared matrices: the stages of one chain
// (column s = stage cs + s of scenario j) or up to kTP crown nodes.  Column arrays in shared memory are
// [element][kTP] (row = 96 bytes, read and written as float4).  Every step is a small non-inlined function so
// that none of them is register-critical for the whole kernel.
// ---------------------------------------------------------------------------------------------------------------
struct SweepSmem {
    float *G, *OT, *L, *B;      // shared matrices (bulk TMA copies of the pack)
    float *X1, *Y, *V, *scr2;   // column arrays: X1 [(nv+nx) or nu rows], Y [max(nv,nu) rows], V [nv rows], scr2 [128 rows]
    int *colnode;               // [kTP] node id of each column
    float *colp;                // [kTP] probability that scales the column's Omega / Theta (Engine.cu:210-221)
    int *anc;                   // [kMaxCs] crown path of the current chain, root first
    uint64_t *mfull;            // [4] mbarriers of the four matrix copies
};
__device__ __forceinline__ SweepSmem sweep_smem(const PArgs &P) {
    SweepSmem S;
    S.G = smem_f(P.oG); S.OT = smem_f(P.oOT); S.L = smem_f(P.oL); S.B = smem_f(P.oB);
    S.X1 = smem_f(P.oX1); S.Y = smem_f(P.oY); S.V = smem_f(P.oV); S.scr2 = smem_f(P.oScr2);
    S.colnode = reinterpret_cast<int *>(smem_f(kOffMisc));
    S.colp = smem_f(kOffMisc + kTP);
    S.anc = reinterpret_cast<int *>(smem_f(kOffMisc + 2 * kTP));
    S.mfull = reinterpret_cast<uint64_t *>(smem_f(kOffBar)) + (kNumBars - 4);
    return S;
}

// one thread: pull G, [OmegaBar | ThetaBar], L, B into the sweep region (it overlays the idle stream ring)
__device__ __forceinline__ void issue_matrix_loads(const PArgs &P) {
    const SweepSmem S = sweep_smem(P);
    fence_proxy_async();   // the region was last written through the generic proxy (vector ring, w, partial sums)
    mbar_expect_tx(&S.mfull[0], P.bG); bulk_g2s(S.G, P.pack + P.pG, P.bG, &S.mfull[0]);
    mbar_expect_tx(&S.mfull[1], P.bOT); bulk_g2s(S.OT, P.pack + P.pOT, P.bOT, &S.mfull[1]);
    mbar_expect_tx(&S.mfull[2], P.bL); bulk_g2s(S.L, P.pack + P.pL, P.bL, &S.mfull[2]);
    mbar_expect_tx(&S.mfull[3], P.bB); bulk_g2s(S.B, P.pack + P.pB, P.bB, &S.mfull[3]);
}

__device__ __forceinline__ void row_load(const float *row, float (&v)[kTP]) {
#pragma unroll
    for (int k = 0; k < kTP / 4; k++) {
        const float4 q = *reinterpret_cast<const float4 *>(row + 4 * k);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
}
__device__ __forceinline__ void row_store(float *row, const float (&v)[kTP]) {
#pragma unroll
    for (int k = 0; k < kTP / 4; k++)
        *reinterpret_cast<float4 *>(row + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// Y[r][c] = sum_k M[r + k*m] X[k][c],  r < m, c < kTP.  M (m x K, column-major), X, Y and scr2 in shared memory.
// 256 threads compute: thread = (rows {rp, rp + 64}, 12 columns, one half of k); the upper half of k is handed over
// through scr2.  All kPC threads call; ends with a CTA barrier.
__device__ __noinline__ void tile_gemm(const float *M, int m, int K, const float *X, float *Y, float *scr2) {
    const int t = threadIdx.x, ks = t >> 7, u = t & 127, rp = u & 63, cg = u >> 6;
    const bool work = ks < 2 && rp < m;
    const bool two = rp + 64 < m;
    float a0[12], a1[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { a0[i] = 0.f; a1[i] = 0.f; }
    if (work) {
        const int kh = (K + 1) >> 1, k0 = ks ? kh : 0, k1 = ks ? K : kh;
        const float *mp = M + (size_t)k0 * m + rp;
        const int d1 = two ? 64 : 0;
        const float *xp = X + k0 * kTP + cg * 12;
#pragma unroll 2
        for (int k = k0; k < k1; k++, mp += m, xp += kTP) {
            const float m0 = mp[0], m1 = mp[d1];
            const float4 x0 = *reinterpret_cast<const float4 *>(xp), x1 = *reinterpret_cast<const float4 *>(xp + 4),
                         x2 = *reinterpret_cast<const float4 *>(xp + 8);
            a0[0] = fmaf(m0, x0.x, a0[0]); a0[1] = fmaf(m0, x0.y, a0[1]); a0[2] = fmaf(m0, x0.z, a0[2]); a0[3] = fmaf(m0, x0.w, a0[3]);
            a0[4] = fmaf(m0, x1.x, a0[4]); a0[5] = fmaf(m0, x1.y, a0[5]); a0[6] = fmaf(m0, x1.z, a0[6]); a0[7] = fmaf(m0, x1.w, a0[7]);
            a0[8] = fmaf(m0, x2.x, a0[8]); a0[9] = fmaf(m0, x2.y, a0[9]); a0[10] = fmaf(m0, x2.z, a0[10]); a0[11] = fmaf(m0, x2.w, a0[11]);
            a1[0] = fmaf(m1, x0.x, a1[0]); a1[1] = fmaf(m1, x0.y, a1[1]); a1[2] = fmaf(m1, x0.z, a1[2]); a1[3] = fmaf(m1, x0.w, a1[3]);
            a1[4] = fmaf(m1, x1.x, a1[4]); a1[5] = fmaf(m1, x1.y, a1[5]); a1[6] = fmaf(m1, x1.z, a1[6]); a1[7] = fmaf(m1, x1.w, a1[7]);
            a1[8] = fmaf(m1, x2.x, a1[8]); a1[9] = fmaf(m1, x2.y, a1[9]); a1[10] = fmaf(m1, x2.z, a1[10]); a1[11] = fmaf(m1, x2.w, a1[11]);
        }
        if (ks == 1) {
            float4 *d = reinterpret_cast<float4 *>(scr2 + u * kTP);
            d[0] = make_float4(a0[0], a0[1], a0[2], a0[3]); d[1] = make_float4(a0[4], a0[5], a0[6], a0[7]);
            d[2] = make_float4(a0[8], a0[9], a0[10], a0[11]); d[3] = make_float4(a1[0], a1[1], a1[2], a1[3]);
            d[4] = make_float4(a1[4], a1[5], a1[6], a1[7]); d[5] = make_float4(a1[8], a1[9], a1[10], a1[11]);
        }
    }
    cbar();
    if (work && ks == 0) {
        const float4 *sp = reinterpret_cast<const float4 *>(scr2 + u * kTP);
        const float4 p0 = sp[0], p1 = sp[1], p2 = sp[2], p3 = sp[3], p4 = sp[4], p5 = sp[5];
        float4 *y0 = reinterpret_cast<float4 *>(Y + rp * kTP + cg * 12);
        y0[0] = make_float4(a0[0] + p0.x, a0[1] + p0.y, a0[2] + p0.z, a0[3] + p0.w);
        y0[1] = make_float4(a0[4] + p1.x, a0[5] + p1.y, a0[6] + p1.z, a0[7] + p1.w);
        y0[2] = make_float4(a0[8] + p2.x, a0[9] + p2.y, a0[10] + p2.z, a0[11] + p2.w);
        if (two) {
            float4 *y1 = reinterpret_cast<float4 *>(Y + (rp + 64) * kTP + cg * 12);
            y1[0] = make_float4(a1[0] + p3.x, a1[1] + p3.y, a1[2] + p3.z, a1[3] + p3.w);
            y1[1] = make_float4(a1[4] + p4.x, a1[5] + p4.y, a1[6] + p4.z, a1[7] + p4.w);
            y1[2] = make_float4(a1[8] + p5.x, a1[9] + p5.y, a1[10] + p5.z, a1[11] + p5.w);
        }
    }
    cbar();
}

__device__ __forceinline__ bool stage_branches(const int *__restrict__ cum, int s) {   // more nodes than the stage above (:699-719)
    return s > 0 && (__ldg(cum + s + 1) - __ldg(cum + s)) > (__ldg(cum + s) - __ldg(cum + s - 1));
}

// the read-only inputs of the forward recursion, read out of P once per step
struct FwdIn { const float *uprev, *uhat_prev, *uhat, *LV; const int *cum; int nu; };
__device__ __forceinline__ FwdIn fwd_in(const PArgs &P) { return FwdIn{P.uprev, P.uhat_prev, P.uhat, P.LV, P.cum, P.nu}; }

// v = ((-1/2 Omega sigma + Theta q_bar) + Psi psi) + Phi xi (:604-627) [df: v = -1/2 Omega r] for every column:
// Y = [OmegaBar | ThetaBar] X1 on entry; Omega_i = OmegaBar / p_i.  V (shared, next GEMM's input) and devVecV.
__device__ __noinline__ void sweep_vcombine(const PArgs &P, int ncols) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, nv = P.nv;
    if (e >= nv) return;
    // every field of P is read into a local first: P lives behind a generic pointer here, so a global store in the
    // loop would otherwise force it to be re-read (same in all the sweep steps below)
    const bool df = P.df_mode != 0;
    const float *__restrict__ p3 = P.part[3], *__restrict__ p2 = P.part[2];
    float *__restrict__ Vg = P.V;
    float y[kTP], b2[kTP], b3[kTP], cp[kTP];
    int cn[kTP];
    row_load(S.Y + e * kTP, y);
    row_load(S.colp, cp);
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
    if (!df) {
#pragma unroll
        for (int s = 0; s < kTP; s++)
            if (s < ncols) {
                const size_t idx = (size_t)cn[s] * nv + e;
                b3[s] = __ldcg(p3 + idx); b2[s] = __ldcg(p2 + idx);
            }
    }
#pragma unroll
    for (int s = 0; s < kTP; s++) {
        float v = 0.f;
        if (s < ncols) {
            v = y[s] / cp[s];
            if (!df) v = (v + b3[s]) + b2[s];
            Vg[(size_t)cn[s] * nv + e] = v;
        }
        y[s] = v;
    }
    row_store(S.V + e * kTP, y);
}

// rows of Y -> dst[node(col)*dim + e]
__device__ __noinline__ void sweep_cols_out(const PArgs &P, int ncols, float *__restrict__ dst, int dim) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x;
    if (e >= dim) return;
    float y[kTP];
    int cn[kTP];
    row_load(S.Y + e * kTP, y);
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
#pragma unroll
    for (int s = 0; s < kTP; s++)
        if (s < ncols) dst[(size_t)cn[s] * dim + e] = y[s];
}

// common end of the backward sweep of a tile: X1 = [-1/2 sigma ; q_bar] (df: -1/2 r) on entry; mpar = parity of the
// matrix mbarriers in this iteration
__device__ __noinline__ void sweep_backward_finish(const PArgs &P, int ncols, uint32_t mpar, int stamp0) {
    const SweepSmem S = sweep_smem(P);
    mbar_wait(&S.mfull[1], mpar);
    tile_gemm(S.OT, P.nv, P.df_mode ? P.nv : P.nv + P.nx, S.X1, S.Y, S.scr2);   // -1/2 OmegaBar sigma + ThetaBar q_bar
    dstamp(P, stamp0);
    sweep_vcombine(P, ncols);
    cbar();
    dstamp(P, stamp0 + 1);
    mbar_wait(&S.mfull[2], mpar);
    tile_gemm(S.L, P.nu, P.nv, S.V, S.Y, S.scr2);                                // L v   (:701, :727)
    dstamp(P, stamp0 + 2);
    sweep_cols_out(P, ncols, P.LV, P.nu);
    cbar();
    dstamp(P, stamp0 + 3);
}

// ---- chains ------------------------------------------------------------------------------------------------------
// columns of chain j: s = 0 .. T-1 <-> node cum[cs + s] + j; also the chain's crown path (root first) for the forward sweep
__device__ __forceinline__ void chain_columns(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, T = P.N - P.cs;
    if (t < kTP) {
        const int node = t < T ? __ldg(P.cum + P.cs + t) + j : 0;
        S.colnode[t] = node;
        S.colp[t] = t < T ? __ldg(P.prob + __ldg(P.omega_idx + node)) : 1.f;
    }
    if (t == 32) {
        int a = __ldg(P.parent + __ldg(P.cum + P.cs) + j);
        for (int k = P.cs - 1; k >= 0; k--) { S.anc[k] = a; a = a >= 0 ? __ldg(P.parent + a) : -1; }
    }
    cbar();
}

// q-scan: q = c + q_child (:651-658).  X1 rows nv.. get q_bar (q of the child, 0 at the leaf); head q -> qh[j]
__device__ __noinline__ void chain_qscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nx = P.nx;
    if (e >= nx) return;
    const float *__restrict__ cg = P.c;
    float *__restrict__ qh = P.qh;
    float *xrow = S.X1 + (P.nv + e) * kTP;
    float cv[kTP];
#pragma unroll
    for (int s = 0; s < kTP; s++) cv[s] = s < T ? __ldcg(cg + (size_t)S.colnode[s] * nx + e) : 0.f;
    float qrun = 0.f;
#pragma unroll
    for (int s = kTP - 1; s >= 0; s--) {
        const float c = cv[s];
        cv[s] = s < T ? qrun : 0.f;
        if (s < T) qrun = c + qrun;
    }
    row_store(xrow, cv);
    qh[(size_t)j * nx + e] = qrun;
}

// r-scan: sigma = beta + r_child (:599); r = ((sigma + D xi) + F psi) + G q_bar (:631-646).  Y = G q_bar on entry.
// X1 rows 0..nv-1 get -1/2 sigma (df: -1/2 r); sigma -> devMatSigma; head r -> rh[j]
__device__ __noinline__ void chain_rscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nv = P.nv;
    if (e >= nv) return;
    const bool df = P.df_mode != 0;
    const float *__restrict__ bg = P.beta, *__restrict__ p0 = P.part[0], *__restrict__ p1 = P.part[1];
    float *__restrict__ sig = P.sigma, *__restrict__ rh = P.rh;
    float y[kTP];
    int cn[kTP];
    row_load(S.Y + e * kTP, y);
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
    float rrun = 0.f;
#pragma unroll
    for (int hb = 1; hb >= 0; hb--) {       // two halves of 12 stages: bounds the registers held by the loads
        float be[12], a0[12], a1[12];
#pragma unroll
        for (int k = 0; k < 12; k++) {
            const int s = hb * 12 + k;
            if (s < T) {
                const size_t idx = (size_t)cn[s] * nv + e;
                be[k] = __ldg(bg + idx); a0[k] = __ldcg(p0 + idx); a1[k] = __ldcg(p1 + idx);
            }
        }
#pragma unroll
        for (int k = 11; k >= 0; k--) {
            const int s = hb * 12 + k;
            float out = 0.f;
            if (s < T) {
                const float sg = be[k] + rrun;
                rrun = ((sg + a0[k]) + a1[k]) + y[s];
                sig[(size_t)cn[s] * nv + e] = sg;
                out = -0.5f * (df ? rrun : sg);
            }
            y[s] = out;
        }
    }
    row_store(S.X1 + e * kTP, y);
    rh[(size_t)j * nv + e] = rrun;
}

__device__ __noinline__ void chain_backward(const PArgs &P, int j, uint32_t mpar) {
    const SweepSmem S = sweep_smem(P);
    chain_columns(P, j);
    chain_qscan(P, j);
    cbar();
    dstamp(P, 3);
    mbar_wait(&S.mfull[0], mpar);
    tile_gemm(S.G, P.nv, P.nx, S.X1 + P.nv * kTP, S.Y, S.scr2);                  // G q_bar   (:644-646)
    dstamp(P, 4);
    chain_rscan(P, j);
    cbar();
    dstamp(P, 5);
    sweep_backward_finish(P, P.N - P.cs, mpar, 6);
}

// u along the crown path root -> node `last` (reference recursion :683-728, element e): returns u of the last node,
// adds every u on the path to usum
__device__ __forceinline__ float path_u(const FwdIn &I, const int *path, int len, int e, float &usum) {
    float up = __ldg(I.uprev + e), uhp = __ldg(I.uhat_prev + e);
    for (int k = 0; k < len; k++) {
        const int a = path[k];
        const float uh = __ldg(I.uhat + (size_t)a * I.nu + e), lv = __ldcg(I.LV + (size_t)a * I.nu + e);
        const float u = stage_branches(I.cum, k) ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;
        usum += u;
        up = u; uhp = uh;
    }
    return up;
}

// u-scan of a chain: u = ((uhat + u_par) - uhat_par) + L v (:722-728); the chain's first stage is a branching stage of
// the reference's loop when it has more nodes than its parent stage (:699-719).  X1 rows 0..nu-1 = u, column T = the
// sum of u over the crown path (for x of the chain's parent)
__device__ __noinline__ void chain_uscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nu = P.nu;
    if (e >= nu) return;
    const FwdIn I = fwd_in(P);
    const int cs = P.cs;
    float *__restrict__ Ug = P.U;
    int cn[kTP];
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
    float uh[kTP], lv[kTP];
#pragma unroll
    for (int s = 0; s < kTP; s++)
        if (s < T) {
            const size_t idx = (size_t)cn[s] * nu + e;
            uh[s] = __ldg(I.uhat + idx); lv[s] = __ldcg(I.LV + idx);
        }
    float usum = 0.f;
    float up = path_u(I, S.anc, cs, e, usum);
    float uhp = cs > 0 ? __ldg(I.uhat + (size_t)S.anc[cs - 1] * nu + e) : __ldg(I.uhat_prev + e);
    const bool head_br = stage_branches(I.cum, cs);
#pragma unroll
    for (int s = 0; s < kTP; s++) {
        float u = 0.f;
        if (s < T) {
            u = (s == 0 && head_br) ? (up + -1.f * uhp) + (uh[s] + lv[s]) : ((uh[s] + up) + -1.f * uhp) + lv[s];
            Ug[(size_t)cn[s] * nu + e] = u;
            up = u; uhp = uh[s];
        } else if (s == T) u = usum;
        lv[s] = u;
    }
    row_store(S.X1 + e * kTP, lv);
}

// x-scan: x = (x_par + e) + B u (:730-737).  Y = B [u | usum] on entry, Y rows = x on exit
__device__ __noinline__ void chain_xscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nx = P.nx;
    if (e >= nx) return;
    const float *__restrict__ eg = P.e;
    float *__restrict__ Xg = P.X;
    const int cs = P.cs;
    const bool head_br = stage_branches(P.cum, cs);
    float y[kTP], ev[kTP];
    int cn[kTP];
    row_load(S.Y + e * kTP, y);
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
#pragma unroll
    for (int s = 0; s < kTP; s++) if (s < T) ev[s] = __ldg(eg + (size_t)cn[s] * nx + e);
    float xrun = __ldg(P.xcur + e);
    for (int k = 0; k < cs; k++) xrun += __ldg(eg + (size_t)S.anc[k] * nx + e);
    if (cs > 0) {
        float bus = 0.f;   // (B usum)[e] sits in column T
#pragma unroll
        for (int s = 0; s < kTP; s++) if (s == T) bus = y[s];
        xrun += bus;
    }
#pragma unroll
    for (int s = 0; s < kTP; s++)
        if (s < T) {
            const float x = (s == 0 && head_br) ? xrun + (ev[s] + y[s]) : (xrun + ev[s]) + y[s];
            Xg[(size_t)cn[s] * nx + e] = x;
            y[s] = x; xrun = x;
        }
    row_store(S.Y + e * kTP, y);
}

// Hx = sysF x, Hu = sysG u (:744-747), t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial sums of
// the two global distances (:792, :810) for every column.  x in Y rows; u in X1 rows (chains) or devVecU (crown).
__device__ __noinline__ void sweep_epilogue(const PArgs &P, int ncols, bool u_in_smem, const float *wxi, const float *wpsi,
                                            double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu;
    const float inv_step = P.inv_step;
    const float *__restrict__ diag = P.diag, *__restrict__ sxmin = P.sxmin, *__restrict__ sxmax = P.sxmax,
                *__restrict__ sxs = P.sxs, *__restrict__ sumin = P.sumin, *__restrict__ sumax = P.sumax;
    const float *Ug = P.U;
    float *__restrict__ pri_xi = P.pri_xi, *__restrict__ pri_psi = P.pri_psi, *__restrict__ dual_xi = P.dual_xi,
          *__restrict__ dual_psi = P.dual_psi;
    double l1 = 0, l2 = 0;
    for (int el = threadIdx.x; el < ny; el += kPC) {
        const bool isx = el < 2 * nx;
        const int jx = el < nx ? el : el - nx, ju = el - 2 * nx;
        const float *src = isx ? S.Y + jx * kTP : S.X1 + ju * kTP;
        for (int sb = 0; sb < ncols; sb += 4) {
            float dgv[4], wv[4], lo[4], hi[4], val[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int s = sb + b;
                if (s < ncols) {
                    const size_t i = (size_t)S.colnode[s];
                    dgv[b] = __ldg(diag + i * ny + el);
                    if (isx) {
                        const size_t kb = i * nx + jx;
                        wv[b] = __ldcg(wxi + i * 2 * nx + el);
                        lo[b] = el < nx ? __ldg(sxmin + kb) : __ldg(sxs + kb);
                        hi[b] = el < nx ? __ldg(sxmax + kb) : __int_as_float(0x7F7F7F7F);
                        val[b] = src[s];
                    } else {
                        const size_t kk = i * nu + ju;
                        wv[b] = __ldcg(wpsi + kk);
                        lo[b] = __ldg(sumin + kk);
                        hi[b] = __ldg(sumax + kk);
                        val[b] = u_in_smem ? src[s] : __ldcg(Ug + kk);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int s = sb + b;
                if (s < ncols) {
                    const size_t i = (size_t)S.colnode[s];
                    const float h = dgv[b] * val[b];
                    const float tt = h + inv_step * wv[b];
                    const float z = clampf(tt, lo[b], hi[b]);
                    if (isx) {
                        pri_xi[i * 2 * nx + el] = h; dual_xi[i * 2 * nx + el] = z;
                        const float df = tt + -1.f * z;
                        if (el < nx) l1 += (double)df * df; else l2 += (double)df * df;
                    } else { pri_psi[i * nu + ju] = h; dual_psi[i * nu + ju] = z; }
                }
            }
        }
    }
    s1 += l1; s2 += l2;
}

__device__ __noinline__ void chain_forward(const PArgs &P, int j, uint32_t mpar, const float *wxi, const float *wpsi,
                                           double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    chain_columns(P, j);
    dstamp(P, 15);
    chain_uscan(P, j);
    cbar();
    dstamp(P, 16);
    mbar_wait(&S.mfull[3], mpar);
    tile_gemm(S.B, P.nx, P.nu, S.X1, S.Y, S.scr2);                               // B u   (:715, :736)
    dstamp(P, 17);
    chain_xscan(P, j);
    cbar();
    dstamp(P, 18);
    sweep_epilogue(P, P.N - P.cs, true, wxi, wpsi, s1, s2);
    cbar();
    dstamp(P, 19);
}

// ---- crown (stages above the chains) ------------------------------------------------------------------------------
// A tile = up to kTP consecutive crown nodes.  solveSumChildren (Utilities.cu:168-201) unrolled over the whole subtree:
//   q_bar_i = sum_{crown j below i} c_j + sum_{heads h below i} q_h
//   sigma_i = beta_i + [ sum_{heads} r_h + sum_{crown j below i} (beta_j + D xi_j + F psi_j) ] + G QS_i
//   QS_i    = sum_{crown j below i} q_bar_j = sum_{crown j below i} (s_j - s_i - 1) c_j + (cs - 1 - s_i) sum_{heads} q_h
// (the descendants of a node are one contiguous id range per stage: children are contiguous, Utilities.cu:184-199)
__device__ __noinline__ void crown_sums(const PArgs &P, int i0, int ncols) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, g = t >> 7, e = t & 127, nx = P.nx, nv = P.nv, cs = P.cs;
    const int head0 = __ldg(P.cum + cs);
    const int *__restrict__ stages = P.stages, *__restrict__ crown_rng = P.crown_rng;
    const float *__restrict__ cg = P.c, *__restrict__ bg = P.beta, *__restrict__ p0 = P.part[0], *__restrict__ p1 = P.part[1],
                *__restrict__ qhg = P.qh, *__restrict__ rhg = P.rh;
    for (int col = 0; col < ncols; col++) {
        const int i = i0 + col, si = __ldg(stages + i);
        const int *rng = crown_rng + (size_t)i * (kMaxCs + 1) * 2;
        float qb = 0.f, qs = 0.f, bs = 0.f;
        for (int s = si + 1; s < cs; s++) {
            const int lo = __ldg(rng + 2 * s), hi = __ldg(rng + 2 * s + 1);
            float cpart = 0.f, bpart = 0.f;
            for (int jn = lo + g; jn < hi; jn += 4) {
                if (e < nx) cpart += __ldcg(cg + (size_t)jn * nx + e);
                if (e < nv) {
                    const size_t idx = (size_t)jn * nv + e;
                    bpart += (__ldg(bg + idx) + __ldcg(p0 + idx)) + __ldcg(p1 + idx);
                }
            }
            qb += cpart; qs += (float)(s - si - 1) * cpart; bs += bpart;
        }
        {
            const int lo = __ldg(rng + 2 * cs) - head0, hi = __ldg(rng + 2 * cs + 1) - head0;
            float hq = 0.f, hr = 0.f;
            for (int h = lo + g; h < hi; h += 4) {
                if (e < nx) hq += __ldcg(qhg + (size_t)h * nx + e);
                if (e < nv) hr += __ldcg(rhg + (size_t)h * nv + e);
            }
            qb += hq; qs += (float)(cs - 1 - si) * hq; bs += hr;
        }
        float *sc = S.scr2 + g * 3 * 128;
        sc[e] = qb; sc[128 + e] = qs; sc[256 + e] = bs;
        cbar();
        if (g == 0) {
            const float *s0 = S.scr2;
            if (e < nx) {
                S.X1[(nv + e) * kTP + col] = ((s0[e] + s0[384 + e]) + s0[768 + e]) + s0[1152 + e];                    // q_bar
                S.V[e * kTP + col] = ((s0[128 + e] + s0[512 + e]) + s0[896 + e]) + s0[1280 + e];                       // QS
            }
            if (e < nv)
                S.X1[e * kTP + col] = __ldg(bg + (size_t)i * nv + e) +
                                      (((s0[256 + e] + s0[640 + e]) + s0[1024 + e]) + s0[1408 + e]);                  // sigma - G QS
        }
        cbar();
    }
    // unused columns: zeros
    if (g == 0)
        for (int col = ncols; col < kTP; col++) {
            if (e < nx) { S.X1[(nv + e) * kTP + col] = 0.f; S.V[e * kTP + col] = 0.f; }
            if (e < nv) S.X1[e * kTP + col] = 0.f;
        }
    cbar();
}

// sigma = (beta + sums) + G QS; X1 rows 0..nv-1: sigma on exit (df: kept unscaled for the second pass)
__device__ __noinline__ void crown_sigma(const PArgs &P, int ncols) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, nv = P.nv;
    if (e >= nv) return;
    const bool df = P.df_mode != 0;
    float *__restrict__ sig = P.sigma;
    float y[kTP], b[kTP];
    int cn[kTP];
    row_load(S.Y + e * kTP, y);
    row_load(S.X1 + e * kTP, b);
#pragma unroll
    for (int s = 0; s < kTP; s++) cn[s] = S.colnode[s];
#pragma unroll
    for (int s = 0; s < kTP; s++) {
        float sg = 0.f;
        if (s < ncols) {
            sg = b[s] + y[s];
            sig[(size_t)cn[s] * nv + e] = sg;
        }
        b[s] = df ? sg : -0.5f * sg;
    }
    row_store(S.X1 + e * kTP, b);
}

// df mode: r = ((sigma + D xi) + F psi) + G q_bar; X1 rows 0..nv-1 = -1/2 r.  Y = G q_bar on entry
__device__ __noinline__ void crown_r_df(const PArgs &P, int ncols) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, nv = P.nv;
    if (e >= nv) return;
    const float *__restrict__ p0 = P.part[0], *__restrict__ p1 = P.part[1];
    float y[kTP], b[kTP];
    row_load(S.Y + e * kTP, y);
    row_load(S.X1 + e * kTP, b);
#pragma unroll
    for (int s = 0; s < kTP; s++) {
        float r = 0.f;
        if (s < ncols) {
            const size_t idx = (size_t)S.colnode[s] * nv + e;
            r = ((b[s] + __ldcg(p0 + idx)) + __ldcg(p1 + idx)) + y[s];
        }
        b[s] = -0.5f * r;
    }
    row_store(S.X1 + e * kTP, b);
}

__device__ __noinline__ void crown_backward(const PArgs &P, int i0, int ncols, uint32_t mpar) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x;
    if (t < kTP) {
        const int node = t < ncols ? i0 + t : 0;
        S.colnode[t] = node;
        S.colp[t] = t < ncols ? __ldg(P.prob + __ldg(P.omega_idx + node)) : 1.f;
    }
    cbar();
    crown_sums(P, i0, ncols);
    dstamp(P, 11);
    mbar_wait(&S.mfull[0], mpar);
    tile_gemm(S.G, P.nv, P.nx, S.V, S.Y, S.scr2);                                // G QS
    crown_sigma(P, ncols);
    cbar();
    if (P.df_mode) {
        tile_gemm(S.G, P.nv, P.nx, S.X1 + P.nv * kTP, S.Y, S.scr2);              // G q_bar
        crown_r_df(P, ncols);
        cbar();
    }
    dstamp(P, 12);
    sweep_backward_finish(P, ncols, mpar, 6);
}

// forward sweep of a crown tile: u along each node's path (reference recursion), x = x_cur + sum_path e + B sum_path u
__device__ __noinline__ void crown_forward(const PArgs &P, int i0, int ncols, uint32_t mpar, const float *wxi, const float *wpsi,
                                           double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, g = t >> 7, e = t & 127, nx = P.nx, nu = P.nu;
    const FwdIn I = fwd_in(P);
    const int *__restrict__ stages = P.stages, *__restrict__ parent = P.parent;
    const float *__restrict__ eg = P.e, *__restrict__ xcur = P.xcur;
    float *__restrict__ Ug = P.U, *__restrict__ Xg = P.X;
    if (t < kTP) S.colnode[t] = t < ncols ? i0 + t : 0;
    cbar();
    for (int col = g; col < kTP; col += 4) {
        float usum = 0.f, xb = 0.f;
        if (col < ncols) {
            const int i = i0 + col, si = __ldg(stages + i);
            int path[kMaxCs];
            int a = i;
#pragma unroll
            for (int k = kMaxCs - 1; k >= 0; k--)
                if (k <= si) { path[k] = a; a = __ldg(parent + a); }
            if (e < nu) {
                const float u = path_u(I, path, si + 1, e, usum);
                Ug[(size_t)i * nu + e] = u;
            }
            if (e < nx) {
                xb = __ldg(xcur + e);
#pragma unroll
                for (int k = 0; k < kMaxCs; k++) if (k <= si) xb += __ldg(eg + (size_t)path[k] * nx + e);
            }
        }
        if (e < nu) S.X1[e * kTP + col] = usum;
        if (e < nx) S.V[e * kTP + col] = xb;
    }
    cbar();
    dstamp(P, 13);
    mbar_wait(&S.mfull[3], mpar);
    tile_gemm(S.B, nx, nu, S.X1, S.Y, S.scr2);                                   // B sum_path u
    if (t < nx) {
        float y[kTP], xb[kTP];
        row_load(S.Y + t * kTP, y);
        row_load(S.V + t * kTP, xb);
#pragma unroll
        for (int s = 0; s < kTP; s++)
            if (s < ncols) {
                const float x = xb[s] + y[s];
                Xg[(size_t)(i0 + s) * nx + t] = x;
                y[s] = x;
            }
        row_store(S.Y + t * kTP, y);
    }
    cbar();
    sweep_epilogue(P, ncols, false, wxi, wpsi, s1, s2);
    cbar();
    dstamp(P, 14);
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
    const int skip = 0;
    fence_proxy_async();   // the ring region was last written through the generic proxy by the sweeps
    long long cyc_empty = 0, cyc_vec = 0, cyc_go = 0;
    const bool pair_ok = (nu & 1) == 0;   // 8-byte copies: every vector starts on an 8-byte boundary when nu is even
    // read everything out of P / R once: the asm statements below clobber memory, so P.x inside the loops would be
    // re-read (through a generic pointer) after every copy
    const int prev_ = (it + 1) & 1;   // W[(it-1)&1] and Y[(it&1)^1]
    const float *g0 = P.pri_xi, *g1 = P.Wxi[prev_], *g2 = P.dual_xi, *g3 = P.Yxi[prev_], *g8 = P.diag;
    const float *g4 = P.pri_psi, *g5 = P.Wpsi[prev_], *g6 = P.dual_psi, *g7 = P.Ypsi[prev_];
    const float *m0 = P.mat[0], *m1 = P.mat[1], *m2 = P.mat[2], *m3 = P.mat[3];
    const int n_mats = P.n_mats, cols_per_chunk = P.cols_per_chunk;
    const int u_begin = R.u_begin, u_end = R.u_end, node_first = R.node_first, node_last = R.node_last;
    auto load_vec = [&](int node, int) {
        const size_t ox = (size_t)node * 2 * nx, op = (size_t)node * nu;
        const long long cv_ = clock64();
        mbar_wait(&M.vempty[vs], vph ^ 1);
        float *dst = M.vec + vs * kVecCount * kVStride;
        const float *s0 = g0 + ox, *s1 = g1 + ox, *s2 = g2 + ox, *s3 = g3 + ox;
        const float *s8 = g8 + (size_t)node * ny;
        const float *s4 = g4 + op, *s5 = g5 + op, *s6 = g6 + op, *s7 = g7 + op;
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
        for (int u = u_begin; u < u_end; u++) {
            const int node = u / n_mats, m = u - node * n_mats;
            if (with_vec && (u == u_begin || m == 0) && node + 2 <= node_last) load_vec(node + 2, it);
            const int ncols = (m & 1) == 0 ? 2 * nx : nu;
            const float *base = (m == 0 ? m0 : (m == 1 ? m1 : (m == 2 ? m2 : m3))) + (size_t)node * nv * ncols;
            for (int c0 = 0; c0 < ncols; c0 += cols_per_chunk) {
                if (skip_n > 0) { skip_n--; continue; }
                if (issued >= limit) return issued;
                const int cc = min(cols_per_chunk, ncols - c0);
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
    if (node_first <= node_last) load_vec(node_first, it);
    if (node_first + 1 <= node_last) load_vec(node_first + 1, it);
    stream_chunks(it, skip, 0x7fffffff, true);
    L.skip = 0;
    L.st = st; L.ph = ph; L.vs = vs; L.vph = vph;
    cyc_go += clock64() - cg_;
    (void)cyc_empty; (void)cyc_vec; (void)cyc_go;
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
    const bool dbg = blockIdx.x == P.clock_cta && threadIdx.x == 0;
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
    // P, R and I live behind generic pointers here: everything the loops need is read into locals first, otherwise
    // each global store below forces those reads to be repeated
    const int n_mats = P.n_mats, u_begin = R.u_begin, u_end = R.u_end, node_last = R.node_last;
    const float inv_step = P.inv_step, step = P.step, a1 = I.a1, a2 = I.a2, sc1 = I.sc1, sc2 = I.sc2;
    const bool br1 = I.br1 != 0, br2 = I.br2 != 0;
    float *__restrict__ Yxi = P.Yxi[I.cur], *__restrict__ Wxi = P.Wxi[I.cur], *__restrict__ Ypsi = P.Ypsi[I.cur],
          *__restrict__ Wpsi = P.Wpsi[I.cur], *__restrict__ cg = P.c;
    float *__restrict__ part0 = P.part[0], *__restrict__ part1 = P.part[1], *__restrict__ part2 = P.part[2],
          *__restrict__ part3 = P.part[3];
    Cand lbx = bx, lbp = bp;
    auto prologue = [&](int node) {
        const long long cp_ = clock64();
        const bool wr_xi = u_begin <= node * n_mats && node * n_mats < u_end;           // owner of (node, D)
        const bool wr_psi = u_begin <= node * n_mats + 1 && node * n_mats + 1 < u_end;   // owner of (node, F)
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
            if (xi_type && (br1 || br2)) {   // distance branch of the previous prox (:792-815; quirk SURVEY A.4-1)
                const float tt = hx + inv_step * wp;
                const float df = tt + -1.f * z;
                if (t < nx) { if (br1) z = z + sc1 * df; }
                else if (br2) { const float d2v = br1 ? (node == 0 ? 0.f : df + -1.f * z) : df; z = z + sc2 * d2v; }
            }
            const float res = hx + -1.f * z;            // computeFixedPointResidual (:839-850)
            const float yn = wp + step * res;           // dualUpdate (:854-864)
            float w = yn * a1;                          // dualExtrapolationStep (:548-552)
            w += a2 * yp;
            wdst[(xi_type ? 0 : kVStride) + t] = w;
            if (xi_type ? wr_xi : wr_psi) {
                const size_t k = xi_type ? (size_t)node * 2 * nx + t : (size_t)node * nu + t;
                if (xi_type) { Yxi[k] = yn; Wxi[k] = w; } else { Ypsi[k] = yn; Wpsi[k] = w; }
                const Cand cd{fabsf(res), res, (int)k};
                if (xi_type) cand_merge(lbx, cd); else cand_merge(lbp, cd);
            }
        }
        ewbar();
        if (wr_xi) {   // c = sysF' xi_w  (:651-658)
            const float *dg = vsl + 8 * kVStride;
            for (int t = et; t < nx; t += kEwWarps * 32) cg[(size_t)node * nx + t] = dg[t] * wdst[t] + dg[nx + t] * wdst[nx + t];
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&M.vempty[vs]); mbar_arrive(&M.wfull[wb]); }
        if (++vs == kVecSlots) { vs = 0; vph ^= 1; }
        if (++wb == 2) { wb = 0; wph ^= 1; }
        cyc_pro += clock64() - cp_;
    };
    if (R.node_first <= node_last) prologue(R.node_first);
    for (int u = u_begin; u < u_end; u++) {
        const int node = u / n_mats, m = u - node * n_mats;
        if ((u == u_begin || m == 0) && node < node_last) prologue(node + 1);   // one node ahead of the GEMV warps
        { const long long c_ = clock64(); mbar_wait(&M.rfull[rb], rph); cyc_rf += clock64() - c_; }
        if (et < nv) {
            const float *rd = M.red + rb * kGemvWarps * kDimMax + et;
            float sum = rd[0];
#pragma unroll
            for (int w = 1; w < kGemvWarps; w++) sum += rd[w * kDimMax];
            float *__restrict__ pm = m == 0 ? part0 : (m == 1 ? part1 : (m == 2 ? part2 : part3));
            pm[(size_t)node * nv + et] = sum;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.rempty[rb]);
        if (++rb == 2) { rb = 0; rph ^= 1; }
    }
    E.vs = vs; E.vph = vph; E.wb = wb; E.wph = wph; E.rb = rb; E.rph = rph;
    bx = lbx; bp = lbp;
    if (blockIdx.x == P.clock_cta && et == 0) { P.phase_ns[30] += cyc_rf; P.phase_ns[31] += cyc_pro; }
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
        for (int s = 0; s < 4; s++) mbar_init(&M.rempty[2 + s], 1);   // = SweepSmem::mfull
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
    if (blockIdx.x == P.clock_cta && tid == 0) g_t_prev = globaltimer();
    auto stamp = [&](int idx) { dstamp(P, idx); };
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
        // the stream ring is idle now: pull the shared sweep matrices over it while the grid barrier is pending
        cbar();
        if (tid == 0) issue_matrix_loads(P);
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
        const uint32_t mpar = (uint32_t)(it & 1);
        // crown tiles: as narrow as the grid allows (every CTA is free during phase C)
        const int tile_w = min(kTP, max(1, (P.n_crown + (int)gridDim.x - 1) / (int)gridDim.x));
        const int n_tiles = (P.n_crown + tile_w - 1) / tile_w;

        // ---- phase B: backward sweep of the chains
        stamp(2);
        for (int j = blockIdx.x; j < P.K; j += gridDim.x) chain_backward(P, j, mpar);
        stamp(10);
        // ---- phase C: backward sweep of the crown (tiles are dealt from the last CTA down: those have the fewest chains)
        if (P.n_crown > 0) {
            grid_sync(P.bar, bar_target);
            stamp(20);
            for (int tl = (int)gridDim.x - 1 - (int)blockIdx.x; tl < n_tiles; tl += gridDim.x)
                crown_backward(P, tl * tile_w, min(tile_w, P.n_crown - tl * tile_w), mpar);
            stamp(21);
            grid_sync(P.bar, bar_target);
            stamp(22);
        }
        // ---- phase F: forward sweep + prox boxes
        const float *wxi = P.Wxi[cur], *wpsi = P.Wpsi[cur];
        for (int tl = (int)gridDim.x - 1 - (int)blockIdx.x; tl < n_tiles; tl += gridDim.x)
            crown_forward(P, tl * tile_w, min(tile_w, P.n_crown - tl * tile_w), mpar, wxi, wpsi, s1, s2);
        for (int j = blockIdx.x; j < P.K; j += gridDim.x) chain_forward(P, j, mpar, wxi, wpsi, s1, s2);
        stamp(23);
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
        stamp(28);
        grid_sync(P.bar, bar_target);
        stamp(29);
    }
    if (blockIdx.x == 0 && tid == 0) *P.iter_dev = P.iters - 1;   // k_finalize finishes iteration iters-1
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct SweepLayout {
    int oG, oOT, oL, oB, oX1, oY, oV, oScr2, end;   // shared-memory float offsets
    int pG, pOT, pL, pB, pack_floats;               // float offsets inside the pack
};
static int pad4(long long n) { return (int)((n + 3) & ~3LL); }
static SweepLayout sweep_layout(const Handle *h) {
    const int nx = h->d.nx, nu = h->d.nu, nv = h->d.nv;
    SweepLayout Y{};
    const int sG = pad4((long long)nv * nx), sOT = pad4((long long)nv * (nv + nx)), sL = pad4((long long)nu * nv), sB = pad4((long long)nx * nu);
    Y.pG = 0; Y.pOT = sG; Y.pL = sG + sOT; Y.pB = sG + sOT + sL; Y.pack_floats = sG + sOT + sL + sB;
    Y.oG = kOffSweep; Y.oOT = Y.oG + sG; Y.oL = Y.oOT + sOT; Y.oB = Y.oL + sL;
    Y.oX1 = Y.oB + sB;
    Y.oY = Y.oX1 + std::max(nv + nx, nu) * kTP;
    Y.oV = Y.oY + std::max(std::max(nv, nu), nx) * kTP;
    Y.oScr2 = Y.oV + nv * kTP;
    Y.end = Y.oScr2 + 128 * kTP;
    return Y;
}
static size_t persist_smem_bytes(const Handle *h) { return (size_t)std::max(kStreamEnd, sweep_layout(h).end) * 4 + 128; }

bool persistent_supported(const Handle *h) {
    const rn_dims &d = h->d;
    if (std::max(2 * d.nx, std::max(d.nu, d.nv)) > kDimMax) return false;
    const int cs = h->chain_stage;
    if (cs >= d.N || cs > kMaxCs) return false;                       // needs a non-branching tail; bounded crown depth
    if (h->h_nps[cs] != d.K) return false;                             // the tail's chains are the scenarios
    if (d.N - cs + (cs > 0 ? 1 : 0) > kTP) return false;               // a chain (+ its parent's column) is one tile
    if (kPStageFloats / d.nv < 1) return false;
    return persist_smem_bytes(h) <= 227 * 1024;
}

rn_status persistent_prepare(Handle *h) {
    if (h->persist_ready) return RN_OK;
    const rn_dims &d = h->d;
    const size_t n = d.nodes;
    const int cs = h->chain_stage, n_crown = h->h_cum[cs];
    RN_CHECK(dev_alloc(h, &h->part[0], n * d.nv)); RN_CHECK(dev_alloc(h, &h->part[1], n * d.nv));
    RN_CHECK(dev_alloc(h, &h->part[2], n * d.nv)); RN_CHECK(dev_alloc(h, &h->part[3], n * d.nv));
    RN_CHECK(dev_alloc(h, &h->LV, n * d.nu));
    RN_CHECK(dev_alloc(h, &h->qh, (size_t)d.K * d.nx)); RN_CHECK(dev_alloc(h, &h->rh, (size_t)d.K * d.nv));
    RN_CHECK(dev_alloc(h, &h->grid_bar, 8));
    RN_CHECK(dev_alloc(h, &h->phase_ns, 32));
    // G | OmegaBar ThetaBar | L | B, each padded to 16 bytes: the four bulk copies of the sweeps
    const SweepLayout Y = sweep_layout(h);
    RN_CHECK(dev_alloc(h, &h->sweep_pack, (size_t)Y.pack_floats));
    const size_t f = sizeof(float);
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pG, h->G, (size_t)d.nv * d.nx * f, cudaMemcpyDeviceToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pOT, h->OmegaBar, (size_t)d.nv * d.nv * f, cudaMemcpyDeviceToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pOT + (size_t)d.nv * d.nv, h->ThetaBar, (size_t)d.nv * d.nx * f, cudaMemcpyDeviceToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pL, h->L, (size_t)d.nu * d.nv * f, cudaMemcpyDeviceToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pB, h->B, (size_t)d.nx * d.nu * f, cudaMemcpyDeviceToDevice, h->stream));
    // descendant id range of every crown node at every later stage up to the chain heads (children are contiguous)
    std::vector<int> rng((size_t)std::max(n_crown, 1) * (kMaxCs + 1) * 2, 0);
    for (int i = 0; i < n_crown; i++) {
        int lo = i, hi = i + 1;
        for (int s = h->h_stages[i] + 1; s <= cs; s++) {
            const int nlo = h->h_child_first[lo], nhi = h->h_child_first[hi - 1] + h->h_child_count[hi - 1];
            lo = nlo; hi = nhi;
            rng[((size_t)i * (kMaxCs + 1) + s) * 2] = lo; rng[((size_t)i * (kMaxCs + 1) + s) * 2 + 1] = hi;
        }
    }
    RN_CHECK(dev_alloc(h, &h->crown_rng, rng.size()));
    RN_CUDA(h, cudaMemcpyAsync(h->crown_rng, rng.data(), rng.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
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
    P.cum = h->cum_dev; P.stages = h->t.stages; P.crown_rng = h->crown_rng;
    P.N = d.N; P.cs = h->chain_stage; P.K = d.K; P.nodes = d.nodes; P.n_crown = h->h_cum[h->chain_stage];
    P.n_mats = h->factor_mode == RN_FACTORS_FULL ? 4 : 2;
    P.df_mode = h->factor_mode == RN_FACTORS_DF ? 1 : 0;
    P.iters = iters; P.nx = d.nx; P.nu = d.nu; P.nv = d.nv;
    P.cols_per_chunk = kPStageFloats / d.nv;
    static const char *clock_env = getenv("RN_CLOCK_CTA");
    P.clock_cta = clock_env ? std::min(atoi(clock_env), h->persist_grid - 1) : 0;
    P.mat[0] = h->D; P.mat[1] = h->F; P.mat[2] = h->Phi; P.mat[3] = h->Psi;
    P.pack = h->sweep_pack; P.diag = h->diag; P.prob = h->t.prob;
    P.beta = h->beta; P.uhat = h->uhat; P.e = h->e; P.xcur = h->xcur; P.uprev = h->uprev; P.uhat_prev = h->uhat_prev;
    P.sxmin = h->sxmin; P.sxmax = h->sxmax; P.sxs = h->sxs; P.sumin = h->sumin; P.sumax = h->sumax;
    P.Yxi[0] = h->yA_xi; P.Yxi[1] = h->yB_xi; P.Ypsi[0] = h->yA_psi; P.Ypsi[1] = h->yB_psi;
    P.Wxi[0] = h->wA_xi; P.Wxi[1] = h->wB_xi; P.Wpsi[0] = h->wA_psi; P.Wpsi[1] = h->wB_psi;
    P.pri_xi = h->pri_xi; P.pri_psi = h->pri_psi; P.dual_xi = h->dual_xi; P.dual_psi = h->dual_psi;
    for (int k = 0; k < 4; k++) P.part[k] = h->part[k];
    P.c = h->c; P.qh = h->qh; P.rh = h->rh; P.sigma = h->sigma; P.V = h->V; P.U = h->U; P.X = h->X; P.LV = h->LV;
    P.dist_part = h->dist_part; P.pinf = h->pinf; P.pinf_part = h->pinf_part;
    P.lambda_tab = h->lambda_tab; P.bar = h->grid_bar; P.iter_dev = h->iter_dev; P.phase_ns = h->phase_ns;
    P.step = h->step; P.inv_step = 1 / h->step; P.pen_x = h->pen_x; P.pen_xs = h->pen_xs;
    const SweepLayout Y = sweep_layout(h);
    P.oG = Y.oG; P.oOT = Y.oOT; P.oL = Y.oL; P.oB = Y.oB; P.oX1 = Y.oX1; P.oY = Y.oY; P.oV = Y.oV; P.oScr2 = Y.oScr2;
    P.pG = Y.pG; P.pOT = Y.pOT; P.pL = Y.pL; P.pB = Y.pB;
    P.bG = (unsigned)(Y.pOT - Y.pG) * 4u; P.bOT = (unsigned)(Y.pL - Y.pOT) * 4u; P.bL = (unsigned)(Y.pB - Y.pL) * 4u;
    P.bB = (unsigned)(Y.pack_floats - Y.pB) * 4u;
    RN_CUDA(h, cudaMemsetAsync(h->grid_bar, 0, sizeof(unsigned int), st));
    RN_CUDA(h, cudaMemsetAsync(h->dist_part, 0, 2 * sizeof(double) * h->persist_grid, st));
    void *args[] = {&P};
    RN_CUDA(h, cudaLaunchCooperativeKernel((const void *)k_apg_persistent, dim3(h->persist_grid), dim3(kPT), args,
                                           persist_smem_bytes(h), st));
    return RN_OK;
}

}  // namespace rn
