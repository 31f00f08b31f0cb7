// rn_persist.cu -- the whole APG loop of SmpcController::algorithmApg as ONE persistent cooperative kernel.
//
// Reference hot loop: /root/reference/src/SmpcController.cu:1500-1525 (about 430 launches per iteration).
// Here every iteration runs inside one resident grid (one CTA per SM, 16 warps) with three software grid barriers per
// iteration (two for a tree without crown, four when the tree is cut across GPUs); DESIGN.md section 3.1:
//
//   phase S  factor stream.  Work unit = (node, matrix in {D, F, Phi, Psi}).  The loader warp keeps a ring of
//            whole-matrix stages full with 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP) of the node's packed Engine
//            factor matrices and fetches the node's dual vectors with cp.async into a 3-deep vector ring, two nodes
//            ahead.  The element-wise warps run the fused pass: finalisation of the PREVIOUS iteration's prox
//            (distance branch), fixed-point residual, dual update y+ = w + step*res, infeasibility log, and the Nesterov
//            extrapolation w = (1+l) y+ - l y of THIS iteration (:535-557, :792-864, :1480-1496) -- the duals are read
//            once and written once per iteration.  The GEMV warps form D xi_w, F psi_w, Phi xi_w, Psi psi_w.
//   sweeps   (:593-747).  The shared matrices G = Bbar', OmegaBar, L and B (every Omega_i is OmegaBar / p_i and
//            Theta q = -1/2 Omega (G q): Engine.cu:707-747) are pulled into shared memory once per iteration by bulk TMA
//            copies that overlay the (then idle) stream ring, so every sweep product is a GEMM
//            [matrix in smem] x [24 columns in smem] across nodes.
//     phase B  chains: below the last branching stage every scenario is an independent chain owned by one CTA; the
//            stage recursion becomes scans (q = c + q_child; sigma = beta + r_child, r = sigma + D xi + F psi +
//            G q_child) around GEMMs across the chain's stages -- no barrier per stage.
//     phase C  crown (stages above the chains): the child->parent sums of solveSumChildren (Utilities.cu:168-201) are
//            unrolled into sums over each node's descendants (contiguous id ranges per stage), so all crown nodes are
//            independent; on one GPU they wait for a counter of published chain heads instead of a grid barrier.
//     phase F  forward sweep (:675-747): u and x are path sums from the root; crown nodes and chains run in the same
//            phase (a chain recomputes its parent's u, x from the crown's L v).  Epilogue: Hx = sysF x, Hu = sysG u,
//            t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial sums of the two global
//            distances of proximalFunG (:792, :810).
//
// Two things shape the code: a 512-thread CTA caps ptxas at 128 registers, and when any function of the
// kernel runs out, EVERY inner loop is scheduled as load -> use pairs -- hence the non-inlined iteration halves with
// their state in local memory; and the sweeps are bound by instruction fetch (32 KB L1.5 instruction cache against
// >200 KB of code per iteration) -- hence compact loops, calls instead of inlined helpers, and per-launch caches.
//
// The last iteration's finalisation is done by k_finalize (rn_apg.cu) after the kernel.
#include <algorithm>
#include <mutex>

#include "rn_internal.h"
#include "rn_device.cuh"

namespace rn {

constexpr int kPC = 512;                        // threads per CTA: 16 warps -> 128 registers per thread
constexpr int kPT = kPC;                        // phase S roles: 11 GEMV warps, 4 element-wise warps, 1 loader warp
constexpr int kPStages = 8;                     // most ring stages (mbarrier slots); the launch picks n_stages x stage_stride
constexpr int kPStageMinFloats = 4096;          // smallest stage payload (16 KB)
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
constexpr int kMaxRanks = 8;                    // GPUs of one node
constexpr double kL2KeepDefault = 0.1;          // RN_L2_KEEP default: share of the L2 given to evict_last factor matrices
constexpr double kL2PrefDefault = 0.0;          // RN_L2_PREFETCH default (share of the L2 refilled with the next iteration's matrices during the
                                                // sweeps): off -- measured on C2/C3/C1r30, phase S does not get faster with up to 18 % L2 hits

struct PArgs {
    const int *parent, *child_first, *child_count, *omega_idx, *cum, *stages;
    const int *pos;                             // [nodes] chain-major row of a node (crown: its id; chain j, stage s: n_crown + j T + s)
    const int *crown_rng;                       // [n_crown][kMaxCs + 1][2]: descendant id range of a crown node per stage
    const int *crown_path;                      // [n_crown][kMaxCs]: path root -> node of every crown node (entry k = its stage-k ancestor)
    unsigned int branch_mask;                   // bit s: stage s has more nodes than stage s-1 (the reference's branching stages, :699-719)
    int N, cs, K, nodes, n_crown, n_mats, df_mode, iters, nx, nu, nv, cols_per_chunk, clock_cta;
    float l2_keep;                              // share of every CTA's stream units whose matrices are loaded with L2 evict_last (0 = plain loads)
    float l2_pref;                              // share of every CTA's stream units (behind the kept ones) pulled into L2 while the sweeps run
    int sh_mode, sh_tile;                       // RN_FACTORS_SHARED: no per-node matrix is read; nodes per tile of its phase S
    int sh_oLt, sh_oC, sh_oGc, sh_oY, sh_oScr2, sh_oVec;   // its shared-memory layout (float offsets; G sits at kOffSweep)
    int n_stages, stage_stride;                 // matrix ring: stages and floats per stage (payload + 32 floats of slack)
    const float *mat[4];                        // D, F, Phi, Psi (packed per node, Engine.cu:201-207)
    const float *pack;                          // G | OmegaBar | L | B, each padded to 16 B (sweeps)
    int nxp, nup, nvp;                          // row lengths of the chain-major arrays (nx, nu, nv rounded up to 4 floats)
    float *cm_c, *cm_lv;                        // chain-major: c = sysF' xi_w [nodes][nxp], L v [nodes][nup]
    const float *cm_beta, *cm_uhat, *cm_e;      // chain-major copies of beta, uhat, e (refreshed at every launch)
    const float *diag, *prob;
    const float *beta, *uhat, *e, *xcur, *uprev, *uhat_prev, *sxmin, *sxmax, *sxs, *sumin, *sumax;
    float *Yxi[2], *Ypsi[2], *Wxi[2], *Wpsi[2];
    float *pri_xi, *pri_psi, *dual_xi, *dual_psi;
    float *part[4];                             // D xi_w, F psi_w, Phi xi_w, Psi psi_w   chain-major [nodes][nvp] each
    float *V, *U, *X;
    double *dist_part;                          // [2*grid]
    // Subtree partition across GPUs (DESIGN.md "multi-GPU"): this rank owns chains [chain_off, chain_off + K) of K_glob;
    // the crown is replicated.  Every rank's exchange buffer is mapped into every process (CUDA IPC); index = rank.
    int n_ranks, rank, K_glob, chain_off;
    unsigned int epoch0;                        // cross-GPU barrier epochs of this launch start after epoch0
    float *hg, *hr;                             // G q / r of this rank's chain heads [K*nv] each (local)
    // Near-root exchange: S_p = sum over the chain heads below bottom-crown node p (stage cs-1) of (G q, r), formed by the rank
    // that owns p's chains in ascending child order (the same bits for any number of ranks) and stored into EVERY rank's
    // table, indexed by crown node id: [n_crown*nvp] each.  solveSumChildren (Utilities.cu:168-201) one level up.
    float *Sg_peer[kMaxRanks], *Sr_peer[kMaxRanks];
    unsigned int *par_ctr;                      // [n_crown] local: chains of a bottom-crown node whose head G q, r are in hg / hr (cumulative)
    unsigned int *s_ctr_peer[kMaxRanks];        // on every rank: bottom-crown nodes whose S row has arrived there (cumulative over launches);
                                                // the owner of a row bumps every rank's counter after storing the row (NVLink atomics)
    unsigned int s_base;                        // value of the counters when this launch starts
    int bottom0, n_bottom, n_owned, own_lo;     // bottom-crown nodes: first id, count; those with their chains on this rank: count, first id (contiguous)
    int n_bottom_active;                        // bottom-crown nodes that have chains on SOME rank (= rows every S table receives per iteration)
    double *dslot_peer[kMaxRanks];              // [kMaxRanks][2] squared prox distances of each rank, on each rank
    unsigned int *xflag_peer[kMaxRanks];        // [kMaxRanks] arrival epochs, on each rank
    int *xerr;                                  // local: set when a cross-GPU wait timed out
    float *pinf4;                               // [iters][4] |res|, res at the arg-max of the xi block and of the psi block
    float *pinf, *pinf_part;                    // [iters], [grid*6]
    const float *lambda_tab;
    unsigned int *bar;
    int *iter_dev;
    unsigned long long *phase_ns;               // [32] fine-grained phase clock of one CTA (see cabi.PHASE_NAMES)
    unsigned long long *cta_ns;                 // [grid][2] per CTA: time spent in phase S (ns, summed over iterations), SM id
    float step, inv_step, pen_x, pen_xs;
    // sweep shared-memory layout (float offsets from the dynamic shared-memory base) and the pack's pieces
    int oG, oOm, oL, oX1, oY, oV, oScr2, oStg, oXb;
    unsigned int bG, bOm, bL, bB, bLt;          // bytes of the bulk copies
    int pG, pOm, pL, pB, pLt;                   // float offsets inside the pack
};

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kPC) : "memory"); }
__device__ __forceinline__ void ewbar() { asm volatile("bar.sync 2, %0;" ::"n"(kEwWarps * 32) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
// fine-grained clock of one CTA (thread 0): accumulates the time since the previous stamp into a shared-memory slot
// (flushed to phase_ns at the end of the kernel; a global read-modify-write per stamp would cost a microsecond each)
__device__ __forceinline__ unsigned long long *clk_smem();
__device__ __noinline__ void dstamp_at(int idx) {
    if (threadIdx.x == 0) {
        unsigned long long *c = clk_smem();
        if (c[33]) {   // this CTA is the clock CTA
            const unsigned long long now = globaltimer();
            c[idx] += now - c[32];
            c[32] = now;
        }
    }
}
__device__ __forceinline__ void dstamp(const PArgs &, int idx) { dstamp_at(idx); }   // a call: the sweeps are bound by instruction fetch
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid barrier over all CTAs (the grid is co-resident: cooperative launch).  Same protocol as
// cooperative_groups::grid_group::sync: CTA barrier, one thread fences + arrives + spins + fences, CTA barrier.
// Every grid barrier of the launch has a fixed ordinal (iteration x barriers per iteration + position), so its arrival
// count is known without per-thread state and any thread can be the one that arrives and polls.
__device__ __forceinline__ unsigned int bar_count(const PArgs &P, int it, int pos) {
    const int per_it = P.n_crown > 0 ? 3 : 2;
    return (unsigned)(it * per_it + pos + 1) * gridDim.x;
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *p >= want (system scope); gives up after about two seconds and raises *err instead of hanging the GPU
__device__ __forceinline__ void wait_sys(const unsigned int *p, unsigned int want, int *err) {
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys_u32(p) - want) < 0) {
        if ((++spins & 1023u) == 0) {
            const unsigned long long now = globaltimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) { *err = 1; break; }
        }
    }
}

// Grid barrier that is also a barrier across the GPUs of the partition: every CTA arrives on the local counter; CTA 0
// waits for all of them, (publish_it >= 0: thread 0 publishes this rank's share of the squared prox distances of
// iteration publish_it to the peers), signals epoch `ep` into every peer's flag array (NVLink stores), waits for every
// peer's signal, then releases the local CTAs.  Peer data written by any thread of this GPU before the barrier is
// ordered before the signal by the CTA barriers + the system-scope fence (cumulativity), like cooperative_groups'
// grid sync does at device scope.
// The barrier comes in two halves so that work that needs no other CTA can sit between them: `cross_arrive` (CTA barrier, this
// CTA's arrival, and on CTA 0 the wait for the local arrivals + the signal to the peers) and `cross_wait` (poll the peers'
// flags, CTA barrier).
__device__ __noinline__ void cross_arrive(const PArgs &P, unsigned int target, unsigned int ep, int publish_it) {
    cbar();
    if (threadIdx.x < 32) {   // one warp; lane r talks to peer r
        const int lane = threadIdx.x;
        if (lane == 0) { __threadfence_system(); atomicAdd(P.bar, 1u); }
        if (blockIdx.x == 0) {
            if (lane == 0) while (ld_acquire_u32(P.bar) < target) {}
            __syncwarp();
            if (publish_it >= 0) {   // this rank's share of the two squared prox distances, CTA order, fixed lane tree
                double t1 = 0, t2 = 0;
                for (int k = lane; k < (int)gridDim.x; k += 32) { t1 += __ldcg(P.dist_part + 2 * k); t2 += __ldcg(P.dist_part + 2 * k + 1); }
                for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o); }
                if (lane < P.n_ranks) {
                    double *ds = P.dslot_peer[lane] + 2 * (publish_it & 1) + 4 * P.rank;
                    ds[0] = t1; ds[1] = t2;
                }
            }
            __threadfence_system();   // lane 0's acquire of the local arrivals + every lane's own stores, before its flag store
            if (lane < P.n_ranks) st_release_sys_u32(P.xflag_peer[lane] + P.rank, ep);
        }
    }
}
__device__ __noinline__ void cross_wait(const PArgs &P, unsigned int ep) {
    if (threadIdx.x < 32) {
        // every CTA polls the arrival flags itself (they live in this GPU's memory): no second hop through a release word
        if ((int)threadIdx.x < P.n_ranks) wait_sys(P.xflag_peer[P.rank] + threadIdx.x, ep, P.xerr);
        __syncwarp();
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
constexpr int kOffMisc = kOffSd + 4;                                    // 4 x kTP words: column -> row / node maps etc.
constexpr int kOffMeta = kOffMisc + 4 * kTP;                             // 64 words: column maps of this CTA's first chain (cache)
constexpr int kOffSegs = kOffMeta + 64;                                 // 64 words: row ranges of the crown node being summed (CrownSegs)
constexpr int kOffFix = kOffSegs + 64;                                  // u_prev | uhat_prev | x_cur, kVStride floats each (per launch)
constexpr int kOffClk = (kOffFix + 3 * kVStride + 1) & ~1;                   // 34 x u64: phase clock accumulators, t_prev, enabled
constexpr int kOffPArgs = kOffClk + 2 * 34;                            // a copy of the kernel arguments (see k_apg_persistent)
constexpr int kPArgsWords = 288;
constexpr int kOffBar = kOffPArgs + kPArgsWords;                                  // mbarriers
constexpr int kNumBars = 2 * kPStages + 2 * kVecSlots + 8 + 8;
constexpr int kOffVec = (kOffBar + 2 * kNumBars + 31) & ~31;            // kVecSlots x kVecCount x kVStride
constexpr int kOffW = kOffVec + kVecSlots * kVecCount * kVStride;       // 2 x kWStride  (xi part | psi part)
constexpr int kOffRed = kOffW + 2 * kWStride;                           // 2 x kGemvWarps x kDimMax
constexpr int kOffRing = kOffRed + 2 * kGemvWarps * kDimMax;            // 128-byte aligned: TMA destination; runtime size
constexpr int kOffSweep = kOffVec;                                      // the sweep region starts where the stream region starts
static_assert(kOffRing % 32 == 0, "ring must be 128-byte aligned");
static_assert(kOffDsh % 2 == 0 && kOffBar % 2 == 0, "8-byte alignment of the double / mbarrier areas");

__device__ __forceinline__ float *smem_f(int off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    return reinterpret_cast<float *>(smem_raw) + off;
}
__device__ __forceinline__ unsigned long long *clk_smem() { return reinterpret_cast<unsigned long long *>(smem_f(kOffClk)); }

// ---------------------------------------------------------------------------------------------------------------
// sweeps.  A "tile" is up to kTP columns that go through the same shared matrices: the stages of one chain
// (column s = stage cs + s of scenario j) or a few crown nodes.  Column arrays in shared memory are [element][kTP]
// (row = 96 bytes, read and written as float4).  The per-node operands of a chain (c, beta, the four partial products,
// uhat, L v, e) live in "chain-major" private arrays -- row pos(node), rows of one chain contiguous, row length padded to
// 16 bytes -- so that ONE bulk TMA copy per array brings a chain's block into the staging area, issued a chain ahead.
// Every step is a small non-inlined function so that none of them is register-critical for the whole kernel; each one
// reads what it needs from P into locals first (P sits behind a generic pointer there: a global store would otherwise
// force a re-read).
// ---------------------------------------------------------------------------------------------------------------
struct SweepSmem {
    float *G, *Om, *L, *B;      // shared matrices (bulk TMA copies of the pack); B overlays G (backward / forward)
    float *X1, *Y, *V, *scr2;   // column arrays: X1 [q_bar rows | sigma rows] or [nu rows], Y [max(nv,nu) rows], V [nv rows]
    float *stg;                 // staging area of the chain blocks
    float *xb;                  // [nx][kTP] crown forward: x_cur + sum_path e per column
    int *colnode;               // [kTP] chain-major row of each column
    int *colid;                 // [kTP] node id of each column
    float *colp;                // [kTP] probability that scales the column's Omega (Engine.cu:210-221)
    int *anc;                   // [kMaxCs] crown path of the current chain, root first
    uint64_t *mfull;            // [4] mbarriers of the matrix copies: G, OmegaBar, L, B
    uint64_t *sfull;            // [4] mbarriers of the staging copies: c | beta, D xi, F psi | Phi xi, Psi psi | uhat, L v, e
};
__device__ __forceinline__ SweepSmem sweep_smem(const PArgs &P) {
    SweepSmem S;
    S.G = smem_f(P.oG); S.Om = smem_f(P.oOm); S.L = smem_f(P.oL); S.B = smem_f(P.oG);
    S.X1 = smem_f(P.oX1); S.Y = smem_f(P.oY); S.V = smem_f(P.oV); S.scr2 = smem_f(P.oScr2); S.stg = smem_f(P.oStg); S.xb = smem_f(P.oXb);
    S.colnode = reinterpret_cast<int *>(smem_f(kOffMisc));
    S.colid = reinterpret_cast<int *>(smem_f(kOffMisc + kTP));
    S.colp = smem_f(kOffMisc + 2 * kTP);
    S.anc = reinterpret_cast<int *>(smem_f(kOffMisc + 3 * kTP));
    S.mfull = reinterpret_cast<uint64_t *>(smem_f(kOffBar)) + (kNumBars - 8);
    S.sfull = S.mfull + 4;
    return S;
}
struct StagePhase { uint32_t c, r, v, f; };   // parities of the staging mbarriers (every thread tracks them)

// one thread: pull OmegaBar, L into the sweep region (it overlays the idle stream ring).  G is not needed: the sweeps take
// G q from the streamed D xi (chain_rscan)
__device__ __forceinline__ void issue_matrix_loads(const PArgs &P) {
    const SweepSmem S = sweep_smem(P);
    fence_proxy_async_all();   // the region was last written through the generic proxy (vector ring, w, partial sums)
    mbar_expect_tx(&S.mfull[1], P.bOm); bulk_g2s(S.Om, P.pack + P.pOm, P.bOm, &S.mfull[1]);
    mbar_expect_tx(&S.mfull[2], P.bL); bulk_g2s(S.L, P.pack + P.pL, P.bL, &S.mfull[2]);
}
// one thread, after the last use of G in this iteration: B takes G's place
__device__ __forceinline__ void issue_b_load(const PArgs &P) {
    const SweepSmem S = sweep_smem(P);
    fence_proxy_async_all();
    mbar_expect_tx(&S.mfull[3], P.bB); bulk_g2s(S.B, P.pack + P.pB, P.bB, &S.mfull[3]);
}
// one thread: the backward operand blocks of chain j -> staging.  Staging layout (floats): beta | p0 | p1 | p2 | p3
__device__ __forceinline__ void issue_chain_backward_loads(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int T = P.N - P.cs;
    const size_t row0 = (size_t)P.n_crown + (size_t)j * T;
    const uint32_t bv = (uint32_t)(T * P.nvp * 4);
    float *d = S.stg;
    fence_proxy_async_all();   // the blocks were written with ordinary stores (phase S of other CTAs, before the barrier)
    mbar_expect_tx(&S.sfull[1], 3 * bv);
    bulk_g2s(d, P.cm_beta + row0 * P.nvp, bv, &S.sfull[1]); d += T * P.nvp;
    bulk_g2s(d, P.part[0] + row0 * P.nvp, bv, &S.sfull[1]); d += T * P.nvp;
    bulk_g2s(d, P.part[1] + row0 * P.nvp, bv, &S.sfull[1]); d += T * P.nvp;
    if (!P.df_mode) {
        mbar_expect_tx(&S.sfull[2], 2 * bv);
        bulk_g2s(d, P.part[2] + row0 * P.nvp, bv, &S.sfull[2]); d += T * P.nvp;
        bulk_g2s(d, P.part[3] + row0 * P.nvp, bv, &S.sfull[2]);
    }
}
// one thread: the forward operand blocks of chain j -> staging.  Layout: uhat | L v | e (T rows each), then the rows of
// the chain's crown path (root first): uhat [cs][nup] | L v [cs][nup] | e [cs][nxp] -- the scans read shared memory only.
// parts & 1: everything that does not depend on the crown of this iteration (the chain's own blocks, the path's uhat and
// e rows) -- issued while the grid barrier before phase F is pending; parts & 2: the path's L v rows (written by the
// crown tiles of other CTAs) and the arrive on the staging barrier.
__device__ __noinline__ void issue_chain_forward_loads(const PArgs &P, int j, int parts) {
    const SweepSmem S = sweep_smem(P);
    const int T = P.N - P.cs, cs = P.cs, nup = P.nup, nxp = P.nxp;
    const size_t row0 = (size_t)P.n_crown + (size_t)j * T;
    const uint32_t bx = (uint32_t)(T * nxp * 4), bu = (uint32_t)(T * nup * 4);
    float *d = S.stg, *dp = S.stg + 2 * T * nup + T * nxp;
    fence_proxy_async_all();
    if (parts & 1) {
        mbar_expect_tx_only(&S.sfull[3], 2 * bu + bx + (uint32_t)cs * (uint32_t)(nup + nxp) * 4u);
        bulk_g2s(d, P.cm_uhat + row0 * nup, bu, &S.sfull[3]); d += T * nup;
        bulk_g2s(d, P.cm_lv + row0 * nup, bu, &S.sfull[3]); d += T * nup;
        bulk_g2s(d, P.cm_e + row0 * nxp, bx, &S.sfull[3]);
    }
    if (parts & 2) mbar_expect_tx(&S.sfull[3], (uint32_t)cs * (uint32_t)nup * 4u);
    if (cs > 0) {
        const int *meta = reinterpret_cast<const int *>(smem_f(kOffMeta));
        const bool hit = meta[0] == j;
        int a = hit ? 0 : __ldg(P.parent + __ldg(P.cum + cs) + j);
#pragma unroll 1
        for (int k = cs - 1; k >= 0; k--) {
            if (hit) a = meta[1 + 2 * kTP + k];
            if (parts & 1) {
                bulk_g2s(dp + k * nup, P.cm_uhat + (size_t)a * nup, (uint32_t)nup * 4u, &S.sfull[3]);
                bulk_g2s(dp + 2 * cs * nup + k * nxp, P.cm_e + (size_t)a * nxp, (uint32_t)nxp * 4u, &S.sfull[3]);
            }
            if (parts & 2) bulk_g2s(dp + (cs + k) * nup, P.cm_lv + (size_t)a * nup, (uint32_t)nup * 4u, &S.sfull[3]);
            if (!hit) a = __ldg(P.parent + a);
        }
    }
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
#ifdef RN_EXP_GEMM_4X12
// Experimental (tools/build_variants.sh, not part of the default build): 4 rows x 12 columns x half of k per thread on 128
// threads -- 16 shared-memory wavefronts per 48 FMAs instead of 14 per 24, one warp per scheduler.  Same summation order
// per element as the default product, so the iterates are bit-identical.
__device__ __noinline__ void tile_gemm(const float *M, int m, int K, const float *X, float *Y, float *scr2) {
    const int t = threadIdx.x, ks = t >> 6, u = t & 63, rp = u & 31, cg = u >> 5;
    const bool work = t < 128 && rp < m;
    const int o1 = rp + 32 < m ? 32 : 0, o2 = rp + 64 < m ? 64 : 0, o3 = rp + 96 < m ? 96 : 0;
    float a[4][12];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 12; i++) a[r][i] = 0.f;
    if (work) {
        const int kh = (K + 1) >> 1, k0 = ks ? kh : 0, k1 = ks ? K : kh;
        const float *mp = M + (size_t)k0 * m + rp;
        const float *xp = X + k0 * kTP + cg * 12;
#pragma unroll 2
        for (int k = k0; k < k1; k++, mp += m, xp += kTP) {
            const float mv[4] = {mp[0], mp[o1], mp[o2], mp[o3]};
            const float4 x0 = *reinterpret_cast<const float4 *>(xp), x1 = *reinterpret_cast<const float4 *>(xp + 4),
                         x2 = *reinterpret_cast<const float4 *>(xp + 8);
            const float xv[12] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w};
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 12; i++) a[r][i] = fmaf(mv[r], xv[i], a[r][i]);
        }
        if (ks == 1) {
            float *d = scr2 + u * 48;
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 12; i++) d[r * 12 + i] = a[r][i];
        }
    }
    cbar();
    if (work && ks == 0) {
        const float *sp = scr2 + u * 48;
        const int off[4] = {0, o1, o2, o3};
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (r == 0 || off[r])
#pragma unroll
                for (int i = 0; i < 12; i++) Y[(rp + off[r]) * kTP + cg * 12 + i] = a[r][i] + sp[r * 12 + i];
    }
    cbar();
}
#elif defined(RN_EXP_GEMM_FFMA2)
// Experimental (tools/build_variants.sh, not part of the default build): the default product with packed FMAs
// (fma.rn.f32x2 = SASS FFMA2: two IEEE fp32 FMAs per instruction, so the results are bit-identical).  The product runs at
// the rate of the FP32 FMA pipe -- 128 x 24 x K lane-FMAs at one 3-register FFMA per two cycles and SM sub-partition is the
// 21 ns per k that the phase clock shows -- so half the FMA instructions should take close to half the variable time.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __noinline__ void tile_gemm(const float *M, int m, int K, const float *X, float *Y, float *scr2) {
    const int t = threadIdx.x, ks = t >> 7, u = t & 127, rp = u & 63, cg = u >> 6;
    const bool work = ks < 2 && rp < m;
    const bool two = rp + 64 < m;
    unsigned long long a0[6], a1[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { a0[i] = 0ull; a1[i] = 0ull; }
    if (work) {
        const int kh = (K + 1) >> 1, k0 = ks ? kh : 0, k1 = ks ? K : kh;
        const float *mp = M + (size_t)k0 * m + rp;
        const int d1 = two ? 64 : 0;
        const float *xp = X + k0 * kTP + cg * 12;
#pragma unroll 2
        for (int k = k0; k < k1; k++, mp += m, xp += kTP) {
            const float m0 = mp[0], m1 = mp[d1];
            const unsigned long long mm0 = pack2(m0, m0), mm1 = pack2(m1, m1);
            const ulonglong2 x0 = *reinterpret_cast<const ulonglong2 *>(xp), x1 = *reinterpret_cast<const ulonglong2 *>(xp + 4),
                             x2 = *reinterpret_cast<const ulonglong2 *>(xp + 8);
            const unsigned long long xv[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
            for (int i = 0; i < 6; i++) { a0[i] = ffma2(mm0, xv[i], a0[i]); a1[i] = ffma2(mm1, xv[i], a1[i]); }
        }
        if (ks == 1) {
            unsigned long long *d = reinterpret_cast<unsigned long long *>(scr2 + u * kTP);
#pragma unroll
            for (int i = 0; i < 6; i++) { d[i] = a0[i]; d[6 + i] = a1[i]; }
        }
    }
    cbar();
    if (work && ks == 0) {
        const float *sp = scr2 + u * kTP;
        float *y0 = Y + rp * kTP + cg * 12, *y1 = Y + (rp + 64) * kTP + cg * 12;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            float lo, hi;
            unpack2(a0[i], lo, hi);
            y0[2 * i] = lo + sp[2 * i]; y0[2 * i + 1] = hi + sp[2 * i + 1];
        }
        if (two) {
#pragma unroll
            for (int i = 0; i < 6; i++) {
                float lo, hi;
                unpack2(a1[i], lo, hi);
                y1[2 * i] = lo + sp[12 + 2 * i]; y1[2 * i + 1] = hi + sp[12 + 2 * i + 1];
            }
        }
    }
    cbar();
}
#else
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
#endif


// The same product for two columns c0, c1 only (a crown tile of one node: the 24-column GEMM would spend 3 us on 1-2
// useful columns).  Thread = (row r = t & 127, half of k = (t >> 7) & 1, column = t >> 8); the halves of k are the ones of
// tile_gemm and are added in the same order, so a crown node gets the same bits whether it is swept alone or inside a wider
// tile -- the crown is replicated across the GPUs of a partition, and the ranks tile it differently.
// All kPC threads call; ends with a CTA barrier.
__device__ __noinline__ void tile_gemv2(const float *M, int m, int K, const float *X, float *Y, float *scr2, int c0, int c1) {
    const int t = threadIdx.x, r = t & 127, h = (t >> 7) & 1, c = t >> 8;
    const int kh = (K + 1) >> 1, k0 = h ? kh : 0, k1 = h ? K : kh;
    const int col = c ? c1 : c0;
    float a = 0.f;
    if (r < m) {
        const float *mp = M + (size_t)k0 * m + r, *x = X + col;
#pragma unroll 4
        for (int k = k0; k < k1; k++, mp += m) a = fmaf(*mp, x[k * kTP], a);
    }
    scr2[h * 256 + c * 128 + r] = a;
    cbar();
    if (h == 0 && r < m) Y[r * kTP + col] = scr2[c * 128 + r] + scr2[256 + c * 128 + r];
    cbar();
}

// all threads: dst[row(col)*ld + e] = src[e][col] for col < ncols, e < dim (rows from `rows`: node ids or chain-major rows)
__device__ __noinline__ void cols_to_global(const float *src, const int *rows, int ncols, int dim, int ld, float *__restrict__ dst) {
    const int e = threadIdx.x & 127;
    if (e >= dim) return;
    src += e * kTP; dst += e;
#pragma unroll 2
    for (int s = threadIdx.x >> 7; s < ncols; s += 4) dst[(size_t)rows[s] * ld] = src[s];
}

__device__ __forceinline__ bool stage_branches(const int *__restrict__ cum, int s) {   // more nodes than the stage above (:699-719)
    return s > 0 && (__ldg(cum + s + 1) - __ldg(cum + s)) > (__ldg(cum + s) - __ldg(cum + s - 1));
}

// v = ((-1/2 Omega sigma + Theta q_bar) + Psi psi) + Phi xi (:604-627), with Theta q_bar = -1/2 Omega (G q_bar) (Theta is
// -1/2 Omega Bbar', Engine.cu:729-734) [df: v = -1/2 Omega r]: Y = OmegaBar X1s on entry, Omega_i = OmegaBar / p_i
// (colp holds 1 / p_i).  All threads: thread = (row, every 4th column).  V rows (next GEMM's input) and devVecV.
__device__ __forceinline__ void sweep_vcombine(const PArgs &P, const SweepSmem &S, int ncols, bool staged) {
    const int e = threadIdx.x & 127, nv = P.nv, nvp = P.nvp, T = P.N - P.cs;
    if (e >= nv) return;
    const bool df = P.df_mode != 0;
    // Phi xi, Psi psi of the columns: staged [col][nvp] (chains) or chain-major global rows (crown tiles: a few columns)
    const float *b2p = staged ? S.stg + 3 * T * nvp + e : P.part[2] + e, *b3p = staged ? b2p + T * nvp : P.part[3] + e;
    float *__restrict__ Vg = P.V + e;
    const float *y = S.Y + e * kTP;
    float *vr = S.V + e * kTP;
#pragma unroll 1
    for (int s = threadIdx.x >> 7; s < kTP; s += 4) {
        float v = 0.f;
        if (s < ncols) {
            float b3 = 0.f, b2 = 0.f;
            if (!df) {
                const size_t row = staged ? (size_t)s : (size_t)S.colnode[s];
                b3 = staged ? b3p[row * nvp] : __ldcg(b3p + row * nvp);
                b2 = staged ? b2p[row * nvp] : __ldcg(b2p + row * nvp);
            }
            v = (y[s] * S.colp[s] + b3) + b2;
            Vg[(size_t)S.colid[s] * nv] = v;
        }
        vr[s] = v;
    }
}

// common end of the backward sweep of a tile: X1s = -1/2 (sigma + G q_bar) (df: -1/2 r) on entry
__device__ __noinline__ void sweep_backward_finish(const PArgs &P, int ncols, bool staged, uint32_t mpar, StagePhase &ph, int next_chain) {
    const SweepSmem S = sweep_smem(P);
    const int nv = P.nv, nu = P.nu, nup = P.nup;
    float *LVg = P.cm_lv;
    const bool one = !staged && ncols == 1;   // a crown tile of one node
    mbar_wait(&S.mfull[1], mpar);
    if (one) tile_gemv2(S.Om, nv, nv, S.X1 + P.nx * kTP, S.Y, S.scr2, 0, 0);
    else tile_gemm(S.Om, nv, nv, S.X1 + P.nx * kTP, S.Y, S.scr2);   // OmegaBar (sigma + G q_bar) (-1/2 folded in)
    dstamp(P, 6);
    if (staged && !P.df_mode) { mbar_wait(&S.sfull[2], ph.v); ph.v ^= 1; }
    sweep_vcombine(P, S, ncols, staged);
    cbar();
    // the staging area is free: request the next chain's blocks, they land while L v is formed and written
    if (next_chain >= 0 && threadIdx.x == 0) issue_chain_backward_loads(P, next_chain);
    dstamp(P, 7);
    mbar_wait(&S.mfull[2], mpar);
    if (one) tile_gemv2(S.L, nu, nv, S.V, S.Y, S.scr2, 0, 0);
    else tile_gemm(S.L, nu, nv, S.V, S.Y, S.scr2);             // L v   (:701, :727)
    dstamp(P, 8);
    cols_to_global(S.Y, S.colnode, ncols, nu, nup, LVg);
    cbar();
    dstamp(P, 9);
}

// ---- chains ------------------------------------------------------------------------------------------------------
// columns of chain j: s = 0 .. T-1 <-> node cum[cs + s] + j; also the chain's crown path (root first) for the forward sweep.
// The maps of the CTA's first chain are kept in shared memory for the whole launch (they cost chains of dependent
// global loads: parent pointers, omega_idx -> prob), so with K <= grid they are computed once.
__device__ __noinline__ void chain_columns(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, T = P.N - P.cs;
    int *meta = reinterpret_cast<int *>(smem_f(kOffMeta));   // [0] chain id, colid[kTP], colp[kTP], anc[kMaxCs]
    static_assert(1 + 2 * kTP + kMaxCs <= 64, "chain metadata cache");
    const bool hit = meta[0] == j;
    if (t < kTP) S.colnode[t] = t < T ? P.n_crown + j * T + t : 0;
    if (hit) {
        if (t < kTP) { S.colid[t] = meta[1 + t]; S.colp[t] = __int_as_float(meta[1 + kTP + t]); }
        if (t < kMaxCs) S.anc[t] = meta[1 + 2 * kTP + t];
        cbar();
        return;
    }
    if (t < kTP) {
        const int node = t < T ? __ldg(P.cum + P.cs + t) + j : 0;
        S.colid[t] = node;
        S.colp[t] = t < T ? 1.f / __ldg(P.prob + __ldg(P.omega_idx + node)) : 1.f;   // Omega_i = OmegaBar / p_i
    }
    if (t == 32) {
        int a = __ldg(P.parent + __ldg(P.cum + P.cs) + j);
        for (int k = P.cs - 1; k >= 0; k--) { S.anc[k] = a; a = a >= 0 ? __ldg(P.parent + a) : -1; }
    }
    cbar();
    if (j == (int)blockIdx.x) {   // every later call is separated from this one by CTA barriers
        if (t < kTP) { meta[1 + t] = S.colid[t]; meta[1 + kTP + t] = __float_as_int(S.colp[t]); }
        if (t < kMaxCs) meta[1 + 2 * kTP + t] = t < P.cs ? S.anc[t] : -1;
        if (t == 0) meta[0] = j;
    }
}

// The scans below are short loops over the chain's stages with scalar shared-memory accesses on purpose: the sweep
// phases are bound by instruction fetch (ncu: "no instruction" is 60-85 % of their stall samples -- every step runs
// once per chain and iteration, and the whole iteration's code does not fit the instruction cache), so a 24-step
// unrolled body costs more in fetch than the bank conflicts of the [element][kTP] column arrays cost here.

// r-scan of a chain.  The reference runs q = c + q_child (:651-658), sigma = beta + r_child (:599) and
// r = ((sigma + D xi) + F psi) + G q_bar (:631-646), q_bar = q of the child.  q itself is never needed, only G q_bar, and
// D_i = G sysF_i' (Engine.cu:720-728) makes G c_i the streamed product D_i xi_w that this scan reads anyway: G q = D xi + G q_child
// runs as a second running sum next to r -- no q-scan and no product with G in the sweeps.
// X1 rows nx.. get -1/2 (sigma + G q_bar) (df: -1/2 r); the head's G q, r -> hg[j], hr[j]
__device__ __noinline__ void chain_rscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nv = P.nv, nvp = P.nvp;
    if (e >= nv) return;
    const bool df = P.df_mode != 0;
    const float *sb = S.stg + e, *s0 = sb + T * nvp, *s1 = s0 + T * nvp;
    float *x1 = S.X1 + (P.nx + e) * kTP;
    for (int s = T; s < kTP; s++) x1[s] = 0.f;
    float rrun = 0.f, grun = 0.f;
#pragma unroll 4
    for (int s = T - 1; s >= 0; s--) {
        const float d = s0[s * nvp];
        const float sg = sb[s * nvp] + rrun;
        rrun = ((sg + d) + s1[s * nvp]) + grun;
        x1[s] = -0.5f * (df ? rrun : sg + grun);
        grun = d + grun;
    }
    P.hg[(size_t)j * nv + e] = grun;
    P.hr[(size_t)j * nv + e] = rrun;
}

__device__ __noinline__ void chain_backward(const PArgs &P, int j, int next_chain, uint32_t mpar, StagePhase &ph) {
    const SweepSmem S = sweep_smem(P);
    chain_columns(P, j);
    dstamp(P, 3);
    mbar_wait(&S.sfull[1], ph.r); ph.r ^= 1;
    chain_rscan(P, j);
    cbar();
    if (threadIdx.x == 0 && P.n_crown > 0) {   // this chain's head G q, r are in the tables: tell whoever sums its parent's heads
        __threadfence();
        atomicAdd(P.par_ctr + S.anc[P.cs - 1], 1u);
    }
    dstamp(P, 5);
    sweep_backward_finish(P, P.N - P.cs, true, mpar, ph, next_chain);
}

// u along the crown path root -> path[len-1] (reference recursion :683-728, element e) for a crown tile: returns u of
// the last node, adds every u on the path to usum.  Crown nodes sit at row = node id of the chain-major arrays.
__device__ __forceinline__ float path_u(const PArgs &P, const int *path, int len, int e, float &usum) {
    const float *__restrict__ uhat = P.cm_uhat + e, *__restrict__ LV = P.cm_lv + e;
    const int nup = P.nup;
    const unsigned bm = P.branch_mask;
    float uh[kMaxCs], lv[kMaxCs];
#pragma unroll
    for (int k = 0; k < kMaxCs; k++)
        if (k < len) { uh[k] = __ldg(uhat + (size_t)path[k] * nup); lv[k] = __ldcg(LV + (size_t)path[k] * nup); }
    const float *fix = smem_f(kOffFix);
    float up = fix[e], uhp = fix[kVStride + e];
#pragma unroll
    for (int k = 0; k < kMaxCs; k++)
        if (k < len) {
            const float u = (bm >> k & 1u) ? (up + -1.f * uhp) + (uh[k] + lv[k]) : ((uh[k] + up) + -1.f * uhp) + lv[k];
            usum += u;
            up = u; uhp = uh[k];
        }
    return up;
}

// u-scan of a chain: u = ((uhat + u_par) - uhat_par) + L v (:722-728), first along the crown path (its rows are staged
// behind the chain's blocks), then down the chain; the chain's first stage is a branching stage of the reference's loop
// when it has more nodes than its parent stage (:699-719).  X1 rows 0..nu-1 = u, column T = the sum of u over the crown
// path (for x of the chain's parent)
__device__ __noinline__ void chain_uscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nu = P.nu, nup = P.nup, cs = P.cs;
    if (e >= nu) return;
    const unsigned bm = P.branch_mask;
    const float *sue = S.stg + e, *sle = sue + T * nup;
    const float *pu = S.stg + 2 * T * nup + T * P.nxp + e, *pl = pu + cs * nup;
    const float *fix = smem_f(kOffFix);
    float up = fix[e], uhp = fix[kVStride + e], usum = 0.f;
#pragma unroll 1
    for (int k = 0; k < cs; k++) {
        const float uh = pu[k * nup], lv = pl[k * nup];
        const float u = (bm >> k & 1u) ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;
        usum += u;
        up = u; uhp = uh;
    }
    const bool head_br = (bm >> cs & 1u) != 0;
    float *x1 = S.X1 + e * kTP;
#pragma unroll 4
    for (int s = 0; s < T; s++) {
        const float uh = sue[s * nup], lv = sle[s * nup];
        const float u = (s == 0 && head_br) ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;
        x1[s] = u;
        up = u; uhp = uh;
    }
    if (T < kTP) x1[T] = usum;
    for (int s = T + 1; s < kTP; s++) x1[s] = 0.f;
}

// x-scan: x = (x_par + e) + B u (:730-737).  Y = B [u | usum] on entry, Y rows = x on exit
__device__ __noinline__ void chain_xscan(const PArgs &P, int j) {
    const SweepSmem S = sweep_smem(P);
    const int e = threadIdx.x, T = P.N - P.cs, nx = P.nx, nxp = P.nxp, cs = P.cs;
    if (e >= nx) return;
    const bool head_br = (P.branch_mask >> cs & 1u) != 0;
    const float *see = S.stg + 2 * T * P.nup + e;
    const float *pe = see + T * nxp + 2 * cs * P.nup;
    float xrun = smem_f(kOffFix)[2 * kVStride + e];
#pragma unroll 1
    for (int k = 0; k < cs; k++) xrun += pe[k * nxp];
    float *y = S.Y + e * kTP;
    if (cs > 0) xrun += y[T];   // (B usum)[e] sits in column T
#pragma unroll 4
    for (int s = 0; s < T; s++) {
        const float ev = see[s * nxp], ys = y[s];
        const float x = (s == 0 && head_br) ? xrun + (ev + ys) : (xrun + ev) + ys;
        y[s] = x; xrun = x;
    }
}

// Hx = sysF x, Hu = sysG u (:744-747), t = Hx + w/step, box projections (Utilities.cu:237-254) and the partial sums of
// the two global distances (:792, :810) for every column.  x in Y rows, u in `urows` rows.  Thread = element of the
// stacked dual [state box | safety level | control box] (so the operand pointers are chosen once), two thread groups
// split the columns, kB columns in flight per trip.
__device__ __noinline__ void sweep_epilogue(const PArgs &P, int ncols, const float *urows, const float *wxi, const float *wpsi,
                                            double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu;
    const int el = threadIdx.x & 255, g = threadIdx.x >> 8;
    static_assert(2 * kDimMax <= 256 && kPC == 512, "one thread per dual element, two groups");
    if (el >= ny) return;
    const int type = el < nx ? 0 : (el < 2 * nx ? 1 : 2);
    const int jx = el - type * nx;                         // index inside the block
    const int bdim = type == 2 ? nu : nx, wdim = type == 2 ? nu : 2 * nx, wo = type == 2 ? jx : el;
    const float *__restrict__ lo_p = (type == 0 ? P.sxmin : (type == 1 ? P.sxs : P.sumin)) + jx;
    const float *__restrict__ hi_p = (type == 2 ? P.sumax : P.sxmax) + jx;
    const float *__restrict__ w_p = (type == 2 ? wpsi : wxi) + wo;
    float *__restrict__ pri = (type == 2 ? P.pri_psi : P.pri_xi) + wo, *__restrict__ dual = (type == 2 ? P.dual_psi : P.dual_xi) + wo;
    const float *__restrict__ diag = P.diag + el;
    const float *vrow = (type == 2 ? urows : S.Y) + jx * kTP;
    const float inv_step = P.inv_step;
    constexpr int kB = 6;
    double l = 0;
#pragma unroll 1
    for (int c0 = g; c0 < ncols; c0 += 2 * kB) {
        float dgv[kB], wv[kB], lo[kB], hi[kB], val[kB];
        int ni[kB];
#pragma unroll
        for (int b = 0; b < kB; b++) {
            const int c = min(c0 + 2 * b, ncols - 1);
            const int i = S.colid[c];
            ni[b] = i;
            dgv[b] = __ldg(diag + (size_t)i * ny);
            wv[b] = __ldcg(w_p + (size_t)i * wdim);
            lo[b] = __ldg(lo_p + (size_t)i * bdim);
            hi[b] = type == 1 ? __int_as_float(0x7F7F7F7F) : __ldg(hi_p + (size_t)i * bdim);
            val[b] = vrow[c];
        }
#pragma unroll
        for (int b = 0; b < kB; b++) {
            if (c0 + 2 * b < ncols) {
                const float h = dgv[b] * val[b];
                const float tt = h + inv_step * wv[b];
                const float z = clampf(tt, lo[b], hi[b]);
                pri[(size_t)ni[b] * wdim] = h; dual[(size_t)ni[b] * wdim] = z;
                const float df = tt + -1.f * z;
                l += (double)df * df;
            }
        }
    }
    if (type == 0) s1 += l; else if (type == 1) s2 += l;
}

__device__ __noinline__ void chain_forward(const PArgs &P, int j, int next_chain, uint32_t mpar, StagePhase &ph, const float *wxi,
                                           const float *wpsi, double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    const int T = P.N - P.cs;
    float *Ug = P.U, *Xg = P.X;
    const int nu = P.nu, nx = P.nx;
    chain_columns(P, j);
    dstamp(P, 15);
    mbar_wait(&S.sfull[3], ph.f); ph.f ^= 1;
    chain_uscan(P, j);
    cbar();
    cols_to_global(S.X1, S.colid, T, nu, nu, Ug);                                 // devVecU
    dstamp(P, 16);
    mbar_wait(&S.mfull[3], mpar);
    tile_gemm(S.B, nx, nu, S.X1, S.Y, S.scr2);                                   // B u   (:715, :736)
    dstamp(P, 17);
    chain_xscan(P, j);
    cbar();
    if (next_chain >= 0 && threadIdx.x == 0) issue_chain_forward_loads(P, next_chain, 3);
    cols_to_global(S.Y, S.colid, T, nx, nx, Xg);                                  // devVecX
    dstamp(P, 18);
    sweep_epilogue(P, T, S.X1, wxi, wpsi, s1, s2);
    cbar();
    dstamp(P, 19);
}

// ---- crown (stages above the chains) ------------------------------------------------------------------------------
// A tile = up to kTP/2 consecutive crown nodes.  solveSumChildren (Utilities.cu:168-201) unrolled over the whole subtree, with
// G c_j = D_j xi_w (see chain_rscan) and g_h = G q_h of a chain head:
//   G q_bar_i = sum_{crown j below i} D xi_j + sum_{heads h below i} g_h
//   G QS_i    = G sum_{crown j below i} q_bar_j = sum_{crown j below i} (s_j - s_i - 1) D xi_j + (cs - 1 - s_i) sum_{heads} g_h
//   sigma_i   = beta_i + [ sum_{heads} r_h + sum_{crown j below i} (beta_j + D xi_j + F psi_j) ] + G QS_i
// (the descendants of a node are one contiguous id range per stage: children are contiguous, Utilities.cu:184-199).
// Sums over rows of chain-major arrays, as float4 columns, for ONE crown node: all row ranges that enter the node's sums --
// its crown descendants stage by stage, then the S rows of the bottom-crown nodes below it -- form one virtual row list, so that
// every load of the node is in flight together (a call per range cost a round trip to L2 each).  The CTA is 8 row groups of two
// warps: the even warp of a pair adds rows of D xi / S_g, the odd warp rows of beta + D xi + F psi / S_r (all nvp long);
// lane = float4 column.  Group g takes virtual rows g, g + 8, ...; kRows of them in flight per trip, added in ascending order.
// acc += row, accw += w(range) * row (even warps: w = how many times the range counts in G QS).
struct F4 { float x, y, z, w; };
struct CrownSegs { int n, start[kMaxCs + 2], lo[kMaxCs + 1]; float w[kMaxCs + 1]; };   // in shared memory; the last range is the S rows
__device__ __noinline__ void crown_row_sums(const PArgs &P, const CrownSegs &G, F4 &acc, F4 &accw) {
    const int t = threadIdx.x, g = t >> 6, role = (t >> 5) & 1, c4 = t & 31;
    constexpr int kRows = 8;
    const int ld = P.nvp;
    if (4 * c4 >= P.nv) return;
    const float *crown0 = (role ? P.cm_beta : P.part[0]) + 4 * c4, *s0 = (role ? P.Sr_peer[P.rank] : P.Sg_peer[P.rank]) + 4 * c4;
    const ptrdiff_t d1 = P.part[0] - P.cm_beta, d2 = P.part[1] - P.cm_beta;
    const int nseg = G.n, total = G.start[nseg];
    F4 sum{0.f, 0.f, 0.f, 0.f}, sumw{0.f, 0.f, 0.f, 0.f};
    int j = 0;
#pragma unroll 1
    for (int v0 = g; v0 < total; v0 += 8 * kRows) {
        float4 v[kRows];
        float wv[kRows];
#pragma unroll
        for (int k = 0; k < kRows; k++) {
            const int vi = v0 + 8 * k;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f); wv[k] = 0.f;
            if (vi < total) {
                while (vi >= G.start[j + 1]) j++;
                const bool is_s = j == nseg - 1;
                const float *q = (is_s ? s0 : crown0) + (size_t)(G.lo[j] + vi - G.start[j]) * ld;
                v[k] = __ldcg(reinterpret_cast<const float4 *>(q));
                wv[k] = G.w[j];
                if (role && !is_s) {
                    const float4 b1 = __ldcg(reinterpret_cast<const float4 *>(q + d1)), b2 = __ldcg(reinterpret_cast<const float4 *>(q + d2));
                    v[k].x = (v[k].x + b1.x) + b2.x; v[k].y = (v[k].y + b1.y) + b2.y; v[k].z = (v[k].z + b1.z) + b2.z; v[k].w = (v[k].w + b1.w) + b2.w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kRows; k++) {
            sum.x += v[k].x; sum.y += v[k].y; sum.z += v[k].z; sum.w += v[k].w;
            sumw.x = fmaf(wv[k], v[k].x, sumw.x); sumw.y = fmaf(wv[k], v[k].y, sumw.y);
            sumw.z = fmaf(wv[k], v[k].z, sumw.z); sumw.w = fmaf(wv[k], v[k].w, sumw.w);
        }
    }
    acc = sum; accw = sumw;
}

// S_p of bottom-crown node p (stage cs-1): G q and r of its chain heads added in ascending child order -- what
// solveSumChildren (Utilities.cu:168-201) leaves in the parent's slot (times G for q) -- stored into every rank's table.  The
// CTA waits for p's chains (a counter they bump after their r-scan).  Threads 0..127: G q rows, 128..255: r rows.
__device__ __noinline__ void parent_sum(const PArgs &P, int p, int it) {
    const int t = threadIdx.x, nv = P.nv;
    const int c0 = __ldg(P.child_first + p) - __ldg(P.cum + P.cs), nc = __ldg(P.child_count + p);   // chain indices of p's heads
    if (t == 0) {
        const unsigned int want = (unsigned)nc * (unsigned)(it + 1);
        while ((int)(ld_acquire_u32(P.par_ctr + p) - want) < 0) {}
        __threadfence();
    }
    cbar();
    const bool isq = t < 128;
    const int e = isq ? t : t - 128, dim = nv;
    if (t < 256 && e < dim) {
        const float *src = (isq ? P.hg : P.hr) + (size_t)c0 * dim + e;
        float acc = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < nc; cb += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = cb + k < nc ? __ldcg(src + (size_t)(cb + k) * dim) : 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) if (cb + k < nc) acc = (cb + k == 0) ? v[k] : acc + v[k];
        }
        const size_t o = (size_t)p * P.nvp + e;
        for (int r = 0; r < P.n_ranks; r++) (isq ? P.Sg_peer[r] : P.Sr_peer[r])[o] = acc;
    }
    cbar();
    if (t < P.n_ranks) {   // the row is in every rank's table: say so on every rank (lane r -> rank r; a peer's counter over NVLink)
        __threadfence_system();
        if (P.n_ranks == 1) atomicAdd(P.s_ctr_peer[0], 1u);
        else atomicAdd_system(P.s_ctr_peer[t], 1u);
    }
}

// `do_wait`: the S rows of ALL bottom-crown nodes -- this rank's and the peers' -- are awaited here (a counter every parent_sum
// bumps on every rank): no barrier across the grid or the GPUs stands between the chains and the crown, which overlaps the rest
// of the chains' backward sweep
__device__ __noinline__ void crown_sums(const PArgs &P, int ncols, bool do_wait, unsigned int wait_s, uint32_t mpar) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, g = t >> 6, role = (t >> 5) & 1, c4 = t & 31, nx = P.nx, nv = P.nv, cs = P.cs, nvp = P.nvp;
    const bool df = P.df_mode != 0;
    static_assert(8 * 3 * 128 <= 128 * kTP, "partial sums of the row groups fit the split-K scratch");
    static_assert(sizeof(CrownSegs) <= 64 * 4, "the range table fits the metadata words behind the chain cache");
    CrownSegs &G = *reinterpret_cast<CrownSegs *>(smem_f(kOffSegs));
    if (do_wait) {
        // while the chains are still scanning: run the GEMM once on whatever X1 holds (two k-steps, result overwritten
        // later) -- it pulls tile_gemm's code into the instruction cache, which is cold at this point of every iteration
        mbar_wait(&S.mfull[1], mpar);
        if (ncols == 1) tile_gemv2(S.Om, nv, 4, S.X1 + nx * kTP, S.Y, S.scr2, 0, 0);
        else tile_gemm(S.Om, nv, 2, S.X1 + nx * kTP, S.Y, S.scr2);
        if (t == 0) {
            if (P.n_ranks == 1) { while ((int)(ld_acquire_u32(P.s_ctr_peer[0]) - wait_s) < 0) {} }
            else wait_sys(P.s_ctr_peer[P.rank], wait_s, P.xerr);
            __threadfence();
        }
        cbar();
    }
#pragma unroll 1
    for (int col = 0; col < ncols; col++) {
        const int i = S.colid[col], si = __ldg(P.stages + i);
        if (t == 0) {   // the node's row ranges: crown descendants of stages si+1 .. cs-1, then the S rows below it (i itself if bottom)
            const int *rng = P.crown_rng + (size_t)i * (kMaxCs + 1) * 2;
            int n = 0, at = 0;
            for (int s = si + 1; s < cs; s++) {
                const int lo = __ldg(rng + 2 * s), hi = __ldg(rng + 2 * s + 1);
                G.start[n] = at; G.lo[n] = lo; G.w[n] = (float)(s - si - 1); at += hi - lo; n++;
            }
            const int blo = si == cs - 1 ? i : __ldg(rng + 2 * (cs - 1)), bhi = si == cs - 1 ? i + 1 : __ldg(rng + 2 * (cs - 1) + 1);
            G.start[n] = at; G.lo[n] = blo; G.w[n] = (float)(cs - 1 - si); at += bhi - blo; n++;
            G.start[n] = at; G.n = n;
        }
        cbar();
        F4 acc{0.f, 0.f, 0.f, 0.f}, accw{0.f, 0.f, 0.f, 0.f};   // even warps: G q_bar, G QS; odd warps: sums of beta + D xi + F psi and of r
        crown_row_sums(P, G, acc, accw);
        {
            float *sc = S.scr2 + (g * 3 + (role ? 2 : 0)) * 128 + 4 * c4;
            *reinterpret_cast<float4 *>(sc) = make_float4(acc.x, acc.y, acc.z, acc.w);
            if (!role) *reinterpret_cast<float4 *>(sc + 128) = make_float4(accw.x, accw.y, accw.z, accw.w);
        }
        cbar();
        if (t < nv) {
            const float *s0 = S.scr2 + t;
            float qb = s0[0], qs = s0[128], bs = s0[256];
#pragma unroll
            for (int k = 1; k < 8; k++) { qb += s0[k * 384]; qs += s0[k * 384 + 128]; bs += s0[k * 384 + 256]; }
            // sigma = (beta + sums) + G QS;  X1 sigma-row = -1/2 (sigma + G q_bar)  (df: -1/2 r, r = ((sigma + D xi) + F psi) + G q_bar)
            const float sg = (__ldg(P.cm_beta + (size_t)i * nvp + t) + bs) + qs;
            S.X1[(nx + t) * kTP + col] = -0.5f * (df ? ((sg + __ldcg(P.part[0] + (size_t)i * nvp + t)) + __ldcg(P.part[1] + (size_t)i * nvp + t)) + qb
                                                     : sg + qb);
        }
        cbar();
    }
    if (t < nv) {   // unused columns: zeros
        float *x1 = S.X1 + (nx + t) * kTP;
        for (int col = ncols; col < kTP; col++) x1[col] = 0.f;
    }
    cbar();
}

// crown nodes in the order of the second crown pass: the upper crown, then the bottom-crown nodes whose chains live on
// OTHER ranks (the owned ones, one contiguous id range, were done in the first pass)
__device__ __forceinline__ int crown_pass2_id(const PArgs &P, int k) { return k < P.own_lo ? k : k + P.n_owned; }

// `mapped`: columns are entries i0 .. i0 + ncols - 1 of the second-pass order; otherwise node ids
__device__ __noinline__ void crown_backward(const PArgs &P, int i0, int ncols, uint32_t mpar, StagePhase &ph, bool do_wait, unsigned int wait_s,
                                            bool mapped) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x;
    if (t < kTP) {
        const int node = t < ncols ? (mapped ? crown_pass2_id(P, i0 + t) : i0 + t) : 0;
        S.colnode[t] = node; S.colid[t] = node;
        S.colp[t] = t < ncols ? 1.f / __ldg(P.prob + __ldg(P.omega_idx + node)) : 1.f;
    }
    cbar();
    crown_sums(P, ncols, do_wait, wait_s, mpar);
    dstamp(P, 11);
    sweep_backward_finish(P, ncols, false, mpar, ph, -1);
}

// forward sweep of a crown tile: u along each node's path (reference recursion), x = x_cur + sum_path e + B sum_path u
__device__ __noinline__ void crown_forward(const PArgs &P, int i0, int ncols, uint32_t mpar, const float *wxi, const float *wpsi,
                                           double &s1, double &s2) {
    const SweepSmem S = sweep_smem(P);
    const int t = threadIdx.x, g = t >> 7, e = t & 127, nx = P.nx, nu = P.nu, nxp = P.nxp;
    const int *__restrict__ stages = P.stages, *__restrict__ cpath = P.crown_path;
    const float *__restrict__ eg = P.cm_e + e;
    float *xbase = S.xb;   // x_cur + sum_path e per column (kept in shared memory: 24 registers live across the GEMM call otherwise)
    if (t < kTP) { S.colnode[t] = t < ncols ? i0 + t : 0; S.colid[t] = S.colnode[t]; }
    cbar();
#pragma unroll 1
    for (int col = g; col < kTP; col += 4) {
        float usum = 0.f, xb = 0.f, u = 0.f;
        if (col < ncols) {
            const int i = i0 + col, si = __ldg(stages + i);
            int path[kMaxCs];
#pragma unroll
            for (int k = 0; k < kMaxCs; k++) path[k] = k <= si ? __ldg(cpath + (size_t)i * kMaxCs + k) : 0;
            if (e < nu) u = path_u(P, path, si + 1, e, usum);
            if (e < nx) {
                float pe[kMaxCs];
#pragma unroll
                for (int k = 0; k < kMaxCs; k++) if (k <= si) pe[k] = __ldg(eg + (size_t)path[k] * nxp);
                xb = smem_f(kOffFix)[2 * kVStride + e];
#pragma unroll
                for (int k = 0; k < kMaxCs; k++) if (k <= si) xb += pe[k];
            }
        }
        if (e < nu) { S.X1[e * kTP + col] = usum; S.V[e * kTP + col] = u; }
        if (e < nx) xbase[e * kTP + col] = xb;
    }
    cbar();
    cols_to_global(S.V, S.colid, ncols, nu, nu, P.U);                             // devVecU
    dstamp(P, 13);
    mbar_wait(&S.mfull[3], mpar);
    if (ncols == 1) tile_gemv2(S.B, nx, nu, S.X1, S.Y, S.scr2, 0, 0);
    else tile_gemm(S.B, nx, nu, S.X1, S.Y, S.scr2);                              // B sum_path u
    if (t < nx) {
        float *y = S.Y + t * kTP;
        const float *xbr = xbase + t * kTP;
#pragma unroll 4
        for (int c = 0; c < kTP; c++) y[c] = c < ncols ? xbr[c] + y[c] : 0.f;
    }
    cbar();
    cols_to_global(S.Y, S.colid, ncols, nx, nx, P.X);                             // devVecX
    sweep_epilogue(P, ncols, S.V, wxi, wpsi, s1, s2);
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
// prefetch != 0: only the first ring-full of chunks of iteration `it` is issued (no vectors) and counted in L.skip --
// called after the CTA's forward sweep, before the closing grid barrier: the factor matrices are constants, so the ring
// fills while the barrier and the next iteration's prologue are pending
__device__ __noinline__ void loader_role(const PArgs &P, const Slice &R, LoaderState &L, int it, int prefetch) {
    const Pipe M = pipe_smem();
    const int nx = P.nx, nu = P.nu, nv = P.nv, lane = threadIdx.x & 31;
    int st = L.st, vs = L.vs; uint32_t ph = L.ph, vph = L.vph;
    const int skip = prefetch ? 0 : L.skip;
    fence_proxy_async();   // the ring region was last written through the generic proxy by the sweeps
    long long cyc_empty = 0, cyc_vec = 0, cyc_go = 0;
    const bool pair_ok = (nu & 1) == 0;   // 8-byte copies: every vector starts on an 8-byte boundary when nu is even
    // read everything out of P / R once: the asm statements below clobber memory, so P.x inside the loops would be
    // re-read (through a generic pointer) after every copy
    const int prev_ = (it + 1) & 1;   // W[(it-1)&1] and Y[(it&1)^1]
    const float *g0 = P.pri_xi, *g1 = P.Wxi[prev_], *g2 = P.dual_xi, *g3 = P.Yxi[prev_];
    const float *g4 = P.pri_psi, *g5 = P.Wpsi[prev_], *g6 = P.dual_psi, *g7 = P.Ypsi[prev_];
    const float *m0 = P.mat[0], *m1 = P.mat[1], *m2 = P.mat[2], *m3 = P.mat[3];
    const int n_mats = P.n_mats, cols_per_chunk = P.cols_per_chunk, n_stages = P.n_stages, stage_stride = P.stage_stride;
    const int u_begin = R.u_begin, u_end = R.u_end, node_first = R.node_first, node_last = R.node_last;
    // L2 residency across iterations: the stream reads the same bytes every iteration, and when they exceed the L2 a plain
    // LRU keeps none of them (cyclic access).  The first u_keep units of the slice are loaded evict_last and the others
    // evict_first, so that share stays in L2 from one iteration to the next and only the rest comes from HBM.
    const int u_keep = P.l2_keep > 0.f ? u_begin + (int)(P.l2_keep * (float)(u_end - u_begin)) : -1;
    auto load_vec = [&](int node, int) {
        const size_t ox = (size_t)node * 2 * nx, op = (size_t)node * nu;
        const long long cv_ = clock64();
        mbar_wait(&M.vempty[vs], vph ^ 1);
        float *dst = M.vec + vs * kVecCount * kVStride;
        const float *s0 = g0 + ox, *s1 = g1 + ox, *s2 = g2 + ox, *s3 = g3 + ox;
        const float *s4 = g4 + op, *s5 = g5 + op, *s6 = g6 + op, *s7 = g7 + op;
        if (pair_ok) {
            for (int k = 2 * lane; k < 2 * nx; k += 64) {
                cp_async8(dst + k, s0 + k); cp_async8(dst + kVStride + k, s1 + k); cp_async8(dst + 2 * kVStride + k, s2 + k);
                cp_async8(dst + 3 * kVStride + k, s3 + k);
            }
            for (int k = 2 * lane; k < nu; k += 64) {
                cp_async8(dst + 4 * kVStride + k, s4 + k); cp_async8(dst + 5 * kVStride + k, s5 + k);
                cp_async8(dst + 6 * kVStride + k, s6 + k); cp_async8(dst + 7 * kVStride + k, s7 + k);
            }
        } else {
            for (int k = lane; k < 2 * nx; k += 32) {
                cp_async4(dst + k, s0 + k); cp_async4(dst + kVStride + k, s1 + k); cp_async4(dst + 2 * kVStride + k, s2 + k);
                cp_async4(dst + 3 * kVStride + k, s3 + k);
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
                    if (u_keep < 0) bulk_g2s(M.ring + st * stage_stride, reinterpret_cast<const void *>(b0), bytes, &M.full[st]);
                    else bulk_g2s_hint(M.ring + st * stage_stride, reinterpret_cast<const void *>(b0), bytes, &M.full[st],
                                       u < u_keep ? kL2EvictLast : kL2EvictFirst);
                }
                __syncwarp();
                if (++st == n_stages) { st = 0; ph ^= 1; }
                issued++;
            }
        }
        return issued;
    };
    if (prefetch) {
        L.skip = stream_chunks(it, 0, n_stages, false);
        L.st = st; L.ph = ph;
        return;
    }
    const long long cg_ = clock64();
    if (node_first <= node_last) load_vec(node_first, it);
    if (node_first + 1 <= node_last) load_vec(node_first + 1, it);
    stream_chunks(it, skip, 0x7fffffff, true);
    L.skip = 0;
    L.st = st; L.ph = ph; L.vs = vs; L.vph = vph;
    cyc_go += clock64() - cg_;
    (void)cyc_empty; (void)cyc_vec; (void)cyc_go;
}

// One thread, at the start of the sweeps: HBM is idle until the next phase S (the sweeps' operands are L2 / shared-memory
// resident), so the matrices this CTA streams right after its kept ones are pulled into L2 now (cp.async.bulk.prefetch.L2,
// SASS UBLKPF.L2): the next phase S finds them there.  Whole nodes [n0, n1) of each of the n_mats arrays, as 16-byte windows.
__device__ __noinline__ void l2_prefetch_next(const PArgs &P, const Slice &R) {
    const int n_units = R.u_end - R.u_begin;
    if (P.l2_pref <= 0.f || n_units <= 0) return;
    const int u0 = R.u_begin + (P.l2_keep > 0.f ? (int)(P.l2_keep * (float)n_units) : 0);
    const int u1 = min(R.u_end, u0 + (int)(P.l2_pref * (float)n_units));
    const int n0 = (u0 + P.n_mats - 1) / P.n_mats, n1 = u1 / P.n_mats;
    if (n1 <= n0) return;
    for (int m = 0; m < P.n_mats; m++) {
        const size_t sz = (size_t)P.nv * ((m & 1) == 0 ? 2 * P.nx : P.nu) * sizeof(float);
        const uintptr_t a0 = (reinterpret_cast<uintptr_t>(P.mat[m]) + (size_t)n0 * sz) & ~uintptr_t(15);
        const uintptr_t a1 = (reinterpret_cast<uintptr_t>(P.mat[m]) + (size_t)n1 * sz + 15) & ~uintptr_t(15);
        for (uintptr_t a = a0; a < a1; a += 65536) bulk_prefetch_l2(reinterpret_cast<const void *>(a), (uint32_t)min((uintptr_t)65536, a1 - a));
    }
}

// ---- GEMV warps: part[m][node] = (factor matrix m of the node) x (w segment).  A warp owns a contiguous block of the
// chunk's columns; lane = rows lane + 32 k of NR row groups, four columns per trip with every load of the trip issued
// before the first FMA (the function is not inlined so that it has its own register budget: inlined into the kernel
// body ptxas serialised load -> FMA pairs through one register and the warps sat on shared-memory latency).
// The nv mod 32 rows that do not fill a row group run in "dot mode" when there are at most 8 of them (RD > 0): lane =
// column, one load + FMA per row and 32 columns, a shuffle reduction per unit -- nv = 97 costs 3 row groups + 1 dot
// row instead of 4 row groups.  Otherwise (RD == 0) the last group over-reads the neighbouring column (or the
// stage's slack): those lanes' sums are never used, so the inner loop carries no row predicate.
template <int NR, int RD>
__device__ __noinline__ void gemv_role(const PArgs &P, const Slice &R, GemvState &G) {
    const Pipe M = pipe_smem();
    constexpr int NA = NR > 0 ? NR : 1, NT = RD > 0 ? RD : 1;
    const int nx = P.nx, nu = P.nu, nv = P.nv, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_stages = P.n_stages, stage_stride = P.stage_stride, n_mats = P.n_mats, cpc = P.cols_per_chunk;
    const int u_begin = R.u_begin, u_end = R.u_end, rem = nv & 31, tail0 = nv - rem;
    // only the 16-byte phase of a chunk's global address matters here (the loader copies 16-byte windows): keep the
    // low four bits of the four array bases in one register
    const unsigned mlow = (unsigned)(reinterpret_cast<uintptr_t>(P.mat[0]) & 15) | (unsigned)(reinterpret_cast<uintptr_t>(P.mat[1]) & 15) << 4 |
                          (unsigned)(reinterpret_cast<uintptr_t>(P.mat[2]) & 15) << 8 | (unsigned)(reinterpret_cast<uintptr_t>(P.mat[3]) & 15) << 12;
    int st = G.st, wb = G.wb, rb = G.rb; uint32_t ph = G.ph, wph = G.wph, rph = G.rph;
    const bool dbg = blockIdx.x == P.clock_cta && threadIdx.x == 0;
    unsigned cyc_full = 0, cyc_w = 0, cyc_red = 0, cyc_cmp = 0;   // low 32 bits of clock differences
    int node = u_begin / n_mats, m = u_begin - node * n_mats;
    bool have_w = false;
    for (int u = u_begin; u < u_end; u++) {
        if (!have_w || m == 0) {
            if (have_w) {   // done with the previous node's w
                __syncwarp();
                if (lane == 0) mbar_arrive(&M.wempty[wb]);
                if (++wb == 2) { wb = 0; wph ^= 1; }
            }
            { const unsigned c_ = (unsigned)clock64(); mbar_wait(&M.wfull[wb], wph); cyc_w += (unsigned)clock64() - c_; }
            have_w = true;
        }
        const bool xi_type = (m & 1) == 0;
        const int ncols = xi_type ? 2 * nx : nu;
        const float *wseg = M.wbuf + wb * kWStride + (xi_type ? 0 : kVStride);
        const unsigned base = ((mlow >> (4 * m)) & 15u) + 4u * (unsigned)node * (unsigned)(nv * ncols);   // mod 16 is all that counts
        float acc[NA], tl[NT];
#pragma unroll
        for (int k = 0; k < NA; k++) acc[k] = 0.f;
#pragma unroll
        for (int r = 0; r < NT; r++) tl[r] = 0.f;
        for (int c0 = 0; c0 < ncols; c0 += cpc) {
            const int cc = min(cpc, ncols - c0);
            const int off = (int)(((base + 4u * (unsigned)(c0 * nv)) & 15u) >> 2);
            const unsigned c1_ = (unsigned)clock64();
            mbar_wait(&M.full[st], ph);
            const unsigned c2_ = (unsigned)clock64();
            cyc_full += c2_ - c1_;
            const float *sb = M.ring + st * stage_stride + off;
            const float *wc = wseg + c0;
            const int ca = cc * warp / kGemvWarps, cb = cc * (warp + 1) / kGemvWarps;   // this warp's columns of the chunk
            if (NR > 0) {
                // volatile shared loads: the compiler keeps their order, so the 4 (NR + 1) loads of a trip are in flight
                // together (written as plain C++ the trip was re-rolled into a load -> FMA chain per column)
                const uint32_t sbase = smem_u32(sb) + 4u * (uint32_t)lane, wbase = smem_u32(wc);
#pragma unroll 1
                for (int j = ca; j < cb; j += 4) {
                    float wv[4], a[4][NA];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int jj = min(j + i, cb - 1);
                        wv[i] = lds_f32(wbase + 4u * (uint32_t)jj);
                        const uint32_t ad = sbase + 4u * (uint32_t)(jj * nv);
#pragma unroll
                        for (int k = 0; k < NA; k++) a[i][k] = lds_f32(ad + 128u * k);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) if (j + i >= cb) wv[i] = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int k = 0; k < NA; k++) acc[k] = fmaf(a[i][k], wv[i], acc[k]);
                }
            }
            if (RD > 0) {   // dot mode: lane = column (a warp's block has at most kDimMax * 2 / kGemvWarps <= 32 columns)
                const int jj = ca + lane;
                if (jj < cb) {
                    const float wl = wc[jj];
                    const float *tp = sb + jj * nv + tail0;
                    float tv[NT];
#pragma unroll
                    for (int r = 0; r < NT; r++) tv[r] = (RD == 1 || r < rem) ? tp[r] : 0.f;
#pragma unroll
                    for (int r = 0; r < NT; r++) tl[r] = fmaf(tv[r], wl, tl[r]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&M.empty[st]);
            if (++st == n_stages) { st = 0; ph ^= 1; }
            cyc_cmp += (unsigned)clock64() - c2_;
        }
        if (RD > 0) {
#pragma unroll
            for (int r = 0; r < NT; r++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tl[r] += __shfl_xor_sync(0xffffffffu, tl[r], o);
        }
        // hand the per-warp partial sums to the element-wise warps
        { const unsigned c_ = (unsigned)clock64(); mbar_wait(&M.rempty[rb], rph ^ 1); cyc_red += (unsigned)clock64() - c_; }
        float *rd = M.red + (rb * kGemvWarps + warp) * kDimMax;
        if (NR > 0) {
#pragma unroll
            for (int k = 0; k < NA; k++) rd[lane + 32 * k] = acc[k];
        }
        if (RD > 0) {
#pragma unroll
            for (int r = 0; r < NT; r++) if (lane == r && r < rem) rd[tail0 + r] = tl[r];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.rfull[rb]);
        if (++rb == 2) { rb = 0; rph ^= 1; }
        if (++m == n_mats) { m = 0; node++; }
    }
    if (have_w) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.wempty[wb]);
        if (++wb == 2) { wb = 0; wph ^= 1; }
    }
    G.st = st; G.ph = ph; G.wb = wb; G.wph = wph; G.rb = rb; G.rph = rph;
    if (dbg) { unsigned long long *c = clk_smem(); c[24] += cyc_full; c[25] += cyc_w; c[26] += cyc_red; c[27] += cyc_cmp; }
}
// row groups / dot rows for a given nv: nv mod 32 in 1..8 -> dot mode
__device__ __forceinline__ void gemv_dispatch(const PArgs &P, const Slice &R, GemvState &G) {
    const int nf = P.nv >> 5, rem = P.nv & 31;
    if (rem == 0 || rem > 8) {
        switch (nf + (rem ? 1 : 0)) {
            case 1: gemv_role<1, 0>(P, R, G); break;
            case 2: gemv_role<2, 0>(P, R, G); break;
            case 3: gemv_role<3, 0>(P, R, G); break;
            default: gemv_role<4, 0>(P, R, G); break;
        }
    } else if (rem == 1) {
        switch (nf) {
            case 0: gemv_role<0, 1>(P, R, G); break;
            case 1: gemv_role<1, 1>(P, R, G); break;
            case 2: gemv_role<2, 1>(P, R, G); break;
            default: gemv_role<3, 1>(P, R, G); break;
        }
    } else {
        switch (nf) {
            case 0: gemv_role<0, 8>(P, R, G); break;
            case 1: gemv_role<1, 8>(P, R, G); break;
            case 2: gemv_role<2, 8>(P, R, G); break;
            default: gemv_role<3, 8>(P, R, G); break;
        }
    }
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
          *__restrict__ Wpsi = P.Wpsi[I.cur];
    const int *__restrict__ pos = P.pos;
    const int nvp = P.nvp;
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
        // (c = sysF' xi_w of :651-658 is not formed: the sweeps take G q from D xi_w, chain_rscan)
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
            pm[(size_t)__ldg(pos + node) * nvp + et] = sum;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&M.rempty[rb]);
        if (++rb == 2) { rb = 0; rph ^= 1; }
    }
    E.vs = vs; E.vph = vph; E.wb = wb; E.wph = wph; E.rb = rb; E.rph = rph;
    bx = lbx; bp = lbp;
    if (blockIdx.x == P.clock_cta && et == 0) { unsigned long long *c = clk_smem(); c[30] += cyc_rf; c[31] += cyc_pro; }
}


// ---------------------------------------------------------------------------------------------------------------
// phase S of the shared-factor formulation (RN_FACTORS_SHARED, "Tier B" of SURVEY.md 8d).  D_i = G sysF_i' and
// F_i = L' sysG_i' (Engine.cu:720-728) with sysF_i = [diag(s_x); diag(s_xs)], sysG_i = diag(s_u), so
//   D_i xi_w = G c_i,  c_i = sysF_i' xi_w      and      F_i psi_w = L' g_i,  g_i = s_u o psi_w
// and no per-node matrix is read at all: the phase is the fused element-wise pass plus two small GEMMs per tile of
// nodes against the SHARED matrices G (nv x nx) and L' (nv x nu), both in shared memory.  A CTA owns a contiguous node
// range and walks it in tiles of up to sh_tile nodes:
//   1. the tile's vectors (Hx, w, z, y of both dual blocks, the diagonals) arrive by nine 1-D bulk TMA copies -- each
//      array's slice of a contiguous node range is contiguous, copied as the enclosing 16-byte window;
//   2. warp = node, lane = element pair: finalisation of the previous prox, residual, dual update, extrapolation
//      (the same arithmetic as ew_role), y and w to global memory, c and g as columns [element][kTP];
//   3. G x [c columns] -> part[0] rows, L' x [g columns] -> part[1] rows (tile_gemm, the sweeps' GEMM).
// The sweeps then run exactly as in RN_FACTORS_DF (v = -1/2 Omega r).
// ---------------------------------------------------------------------------------------------------------------
struct ShState { uint32_t vph; };
struct ShRange { int n0, n1; };
__device__ __forceinline__ ShRange sh_range(const PArgs &P) {
    ShRange R;
    R.n0 = (int)((long long)P.nodes * blockIdx.x / gridDim.x);
    R.n1 = (int)((long long)P.nodes * (blockIdx.x + 1) / gridDim.x);
    return R;
}
__device__ __forceinline__ uint64_t *sh_bar(int k) { return reinterpret_cast<uint64_t *>(smem_f(kOffBar)) + k; }   // Pipe::full[k]: idle in this mode
// one thread: G and L' of iteration `it` -> shared memory (barrier 0, parity it & 1).  Issued while the closing grid
// barrier of the previous iteration is pending (the sweep region is dead by then; the matrices are constants).
__device__ __noinline__ void sh_issue_matrices(const PArgs &P) {
    fence_proxy_async_all();
    mbar_expect_tx(sh_bar(0), P.bG + P.bLt);
    bulk_g2s(smem_f(kOffSweep), P.pack + P.pG, P.bG, sh_bar(0));
    bulk_g2s(smem_f(P.sh_oLt), P.pack + P.pLt, P.bLt, sh_bar(0));
}
struct ShVecPtrs { const float *g[9]; };
// global sources of the nine vector arrays of iteration `it` (same roles as loader_role's g0..g8)
__device__ __forceinline__ ShVecPtrs sh_vec_ptrs(const PArgs &P, int it) {
    const int prev_ = (it + 1) & 1;   // W[(it-1)&1] and Y[(it&1)^1]
    ShVecPtrs V;
    V.g[0] = P.pri_xi; V.g[1] = P.Wxi[prev_]; V.g[2] = P.dual_xi; V.g[3] = P.Yxi[prev_];
    V.g[4] = P.pri_psi; V.g[5] = P.Wpsi[prev_]; V.g[6] = P.dual_psi; V.g[7] = P.Ypsi[prev_]; V.g[8] = P.diag;
    return V;
}
// float offset of array a's buffer inside the vector buffer / its row length
__device__ __forceinline__ int sh_vec_off(const PArgs &P, int a) {
    const int tile = P.sh_tile, bx = ((tile * 2 * P.nx + 3) & ~3) + 8, bp = ((tile * P.nu + 3) & ~3) + 8;
    return a < 4 ? a * bx : (4 * bx + (a - 4) * bp);
}
__device__ __forceinline__ int sh_vec_dim(const PArgs &P, int a) { return a < 4 ? 2 * P.nx : (a < 8 ? P.nu : 2 * P.nx + P.nu); }
// one thread: the vectors of nodes [t0, t0 + nt) -> the vector buffer (barrier 1)
__device__ __noinline__ void sh_issue_vectors(const PArgs &P, int it, int t0, int nt) {
    const ShVecPtrs V = sh_vec_ptrs(P, it);
    fence_proxy_async_all();
    uint32_t total = 0;
    uintptr_t src[9]; uint32_t bytes[9];
#pragma unroll
    for (int a = 0; a < 9; a++) {
        const int dim = sh_vec_dim(P, a);
        const uintptr_t p0 = reinterpret_cast<uintptr_t>(V.g[a] + (size_t)t0 * dim), p1 = p0 + (size_t)nt * dim * sizeof(float);
        src[a] = p0 & ~uintptr_t(15);
        bytes[a] = (uint32_t)(((p1 + 15) & ~uintptr_t(15)) - src[a]);
        total += bytes[a];
    }
    mbar_expect_tx(sh_bar(1), total);
#pragma unroll
    for (int a = 0; a < 9; a++) bulk_g2s(smem_f(P.sh_oVec + sh_vec_off(P, a)), reinterpret_cast<const void *>(src[a]), bytes[a], sh_bar(1));
}

__device__ __noinline__ void sh_elementwise(const PArgs &P, const EwIter &I, int it, int t0, int nt, Cand &bx, Cand &bp) {
    const int nx = P.nx, nu = P.nu, ny = 2 * nx + nu, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_step = P.inv_step, step = P.step, a1 = I.a1, a2 = I.a2, sc1 = I.sc1, sc2 = I.sc2;
    const bool br1 = I.br1 != 0, br2 = I.br2 != 0;
    float *__restrict__ Yxi = P.Yxi[I.cur], *__restrict__ Wxi = P.Wxi[I.cur], *__restrict__ Ypsi = P.Ypsi[I.cur],
          *__restrict__ Wpsi = P.Wpsi[I.cur];
    float *ccol = smem_f(P.sh_oC), *gcol = smem_f(P.sh_oGc);
    const int *colnode = reinterpret_cast<const int *>(smem_f(kOffMisc));
    // every array's slice starts `skip` floats into its 16-byte window
    const ShVecPtrs V = sh_vec_ptrs(P, it);
    const float *b[9];
#pragma unroll
    for (int a = 0; a < 9; a++) {
        const uintptr_t p0 = reinterpret_cast<uintptr_t>(V.g[a] + (size_t)t0 * sh_vec_dim(P, a));
        b[a] = smem_f(P.sh_oVec + sh_vec_off(P, a)) + (int)((p0 & 15) >> 2);
    }
    Cand lbx = bx, lbp = bp;
#pragma unroll 1
    for (int nl = warp; nl < kTP; nl += kPC / 32) {
        if (nl >= nt) {   // unused columns of the tile: zeros
            for (int t = lane; t < nx; t += 32) ccol[t * kTP + nl] = 0.f;
            for (int t = lane; t < nu; t += 32) gcol[t * kTP + nl] = 0.f;
            continue;
        }
        const int node = t0 + nl;
        const float *hx = b[0] + nl * 2 * nx, *wp = b[1] + nl * 2 * nx, *zz = b[2] + nl * 2 * nx, *yp = b[3] + nl * 2 * nx;
        const float *dg = b[8] + nl * ny;
        // state-box element t and safety element nx + t of the node: together they give c[t] = (sysF' xi_w)[t]  (:651-658)
        for (int t = lane; t < nx; t += 32) {
            float wv[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int el = t + h * nx;
                const float hxv = hx[el], wpv = wp[el], ypv = yp[el];
                float z = zz[el];
                if (br1 || br2) {   // distance branch of the previous prox (:792-815; quirk SURVEY A.4-1)
                    const float tt = hxv + inv_step * wpv;
                    const float df = tt + -1.f * z;
                    if (h == 0) { if (br1) z = z + sc1 * df; }
                    else if (br2) { const float d2v = br1 ? (node == 0 ? 0.f : df + -1.f * z) : df; z = z + sc2 * d2v; }
                }
                const float res = hxv + -1.f * z;           // computeFixedPointResidual (:839-850)
                const float yn = wpv + step * res;          // dualUpdate (:854-864)
                float w = yn * a1;                          // dualExtrapolationStep (:548-552)
                w += a2 * ypv;
                const size_t k = (size_t)node * 2 * nx + el;
                Yxi[k] = yn; Wxi[k] = w;
                cand_merge(lbx, Cand{fabsf(res), res, (int)k});
                wv[h] = w;
            }
            const float cv = dg[t] * wv[0] + dg[nx + t] * wv[1];
            ccol[t * kTP + nl] = cv;
        }
        const float *hp = b[4] + nl * nu, *wpp = b[5] + nl * nu, *zp = b[6] + nl * nu, *ypp = b[7] + nl * nu;
        for (int t = lane; t < nu; t += 32) {
            const float hxv = hp[t], wpv = wpp[t], ypv = ypp[t], z = zp[t];
            const float res = hxv + -1.f * z;
            const float yn = wpv + step * res;
            float w = yn * a1;
            w += a2 * ypv;
            const size_t k = (size_t)node * nu + t;
            Ypsi[k] = yn; Wpsi[k] = w;
            cand_merge(lbp, Cand{fabsf(res), res, (int)k});
            gcol[t * kTP + nl] = dg[2 * nx + t] * w;       // g = sysG' psi_w
        }
    }
    bx = lbx; bp = lbp;
}

__device__ __noinline__ void sh_stream(const PArgs &P, ShState &S, const EwIter &I, int it, Cand &bx, Cand &bp) {
    const ShRange R = sh_range(P);
    const int tile = P.sh_tile, nv = P.nv, nvp = P.nvp;
    int *colnode = reinterpret_cast<int *>(smem_f(kOffMisc));
    float *Ysm = smem_f(P.sh_oY), *scr2 = smem_f(P.sh_oScr2);
    if (threadIdx.x == 0 && R.n0 < R.n1) sh_issue_vectors(P, it, R.n0, min(tile, R.n1 - R.n0));
    bool mats = false;
#pragma unroll 1
    for (int t0 = R.n0; t0 < R.n1; t0 += tile) {
        const int nt = min(tile, R.n1 - t0);
        if (threadIdx.x < kTP) colnode[threadIdx.x] = threadIdx.x < nt ? __ldg(P.pos + t0 + threadIdx.x) : 0;
        mbar_wait(sh_bar(1), S.vph); S.vph ^= 1;
        cbar();
        sh_elementwise(P, I, it, t0, nt, bx, bp);
        cbar();
        // the vector buffer is free: the next tile's vectors land while this tile's products are formed
        if (threadIdx.x == 0 && t0 + tile < R.n1) sh_issue_vectors(P, it, t0 + tile, min(tile, R.n1 - t0 - tile));
        if (!mats) { mbar_wait(sh_bar(0), (uint32_t)(it & 1)); mats = true; }
        tile_gemm(smem_f(kOffSweep), nv, P.nx, smem_f(P.sh_oC), Ysm, scr2);          // G c = D xi_w
        cols_to_global(Ysm, colnode, nt, nv, nvp, P.part[0]);
        tile_gemm(smem_f(P.sh_oLt), nv, P.nu, smem_f(P.sh_oGc), Ysm, scr2);          // L' g = F psi_w
        cols_to_global(Ysm, colnode, nt, nv, nvp, P.part[1]);
        cbar();
    }
    if (!mats) mbar_wait(sh_bar(0), (uint32_t)(it & 1));   // keep the barrier's phase in step on a CTA without nodes
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
// Per-thread state that lives across iterations.  It sits in local memory and the two halves of an iteration are
// non-inlined functions that get it by reference: the kernel body keeps almost nothing live across the calls, so every
// role / sweep function below has (nearly) the whole 128-register budget of a 512-thread CTA.  (With the iteration
// inlined into the kernel, ptxas ran out of registers and scheduled every inner loop as load -> use pairs.)
struct KState {
    Slice R;
    StagePhase SP;
    GemvState GS;
    EwState ES;
    LoaderState LS;
    ShState SH;
    double s1, s2;
};

// infeasibility log entry of iteration it-1 from the per-CTA candidates (one warp; every load issued before the merge)
__device__ __noinline__ void pinf_merge(const PArgs &P, int it) {
    const int lane = threadIdx.x & 31, grid = (int)gridDim.x;
    Cand x{-1.f, 0.f, 0x7fffffff}, p{-1.f, 0.f, 0x7fffffff};
    for (int b0 = 0; b0 < grid; b0 += 128) {
        float2 v[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int b = min(b0 + lane + 32 * k, grid - 1);
            const float2 *o = reinterpret_cast<const float2 *>(P.pinf_part + 6 * (size_t)b);   // 24-byte records
            v[k][0] = __ldcg(o); v[k][1] = __ldcg(o + 1); v[k][2] = __ldcg(o + 2);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {   // a clamped duplicate merges to itself
            cand_merge(x, Cand{v[k][0].x, v[k][0].y, __float_as_int(v[k][1].x)});
            cand_merge(p, Cand{v[k][1].y, v[k][2].x, __float_as_int(v[k][2].y)});
        }
    }
    x = cand_warp(x); p = cand_warp(p);
    if (lane == 0) {
        P.pinf[it - 1] = fmaxf(x.v, p.v);    // max(maxValueXi, maxValuePsi) (:1495)
        float *o4 = P.pinf4 + 4 * (size_t)(it - 1);
        o4[0] = x.a; o4[1] = x.v; o4[2] = p.a; o4[3] = p.v;   // for the cross-rank merge on the host
    }
}

__device__ __noinline__ bool cta_sweeps_backward(const PArgs &P);
// first half of iteration `it`: global distances of the previous prox, phase S, infeasibility log of iteration it-1
__device__ __noinline__ void iter_stream(const PArgs &P, KState &K, int it) {
    double *dsh = reinterpret_cast<double *>(smem_f(kOffDsh));      // 2 * 16 doubles
    Cand *csh = reinterpret_cast<Cand *>(smem_f(kOffCsh));          // 2 * 16 candidates
    float *sd = smem_f(kOffSd);                                     // d1, d2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long t_in = tid == 0 ? globaltimer() : 0ull;
    // ---- global distances of the previous iteration's prox (cublasSnrm2, :792, :810); zeros at it == 0
    if (P.n_ranks > 1) {   // every rank published its share (crown counted by rank 0 only); same order everywhere
        if (tid == 0) {
            double t1 = 0, t2 = 0;
            const double *ds = P.dslot_peer[P.rank] + 2 * ((it + 1) & 1);
            if (it > 0) for (int r = 0; r < P.n_ranks; r++) { t1 += __ldcg(ds + 4 * r); t2 += __ldcg(ds + 4 * r + 1); }
            sd[0] = (float)sqrt(t1); sd[1] = (float)sqrt(t2);
        }
        cbar();
    }   // one GPU: the loader warp left them in sd[] when it closed the previous iteration's barrier (iter_close)
    Cand bx{-1.f, 0.f, 0x7fffffff}, bp{-1.f, 0.f, 0x7fffffff};

    // ---- phase S: three-role pipeline.  loader -> [ring] -> GEMV warps -> [red] -> element-wise warps -> [wbuf] -> GEMV
    if (P.sh_mode) {   // shared-factor formulation: no matrix stream, every warp runs the element-wise pass
        const float lam = __ldg(P.lambda_tab + it);
        const float d1 = sd[0], d2 = sd[1];
        const float thr1 = P.inv_step * P.pen_x, thr2 = P.inv_step * P.pen_xs;
        EwIter I;
        I.a1 = 1.f + lam; I.a2 = -lam; I.cur = it & 1;
        I.br1 = d1 > thr1; I.br2 = d2 > thr2;
        I.sc1 = I.br1 ? 1.f - thr1 / d1 : 0.f; I.sc2 = I.br2 ? 1.f - thr2 / d2 : 0.f;
        sh_stream(P, K.SH, I, it, bx, bp);
    } else if (warp < kGemvWarps) {
        gemv_dispatch(P, K.R, K.GS);
    } else if (warp == kLoaderWarp) {
        loader_role(P, K.R, K.LS, it, 0);
    } else {
        const float lam = __ldg(P.lambda_tab + it);
        const float d1 = sd[0], d2 = sd[1];
        const float thr1 = P.inv_step * P.pen_x, thr2 = P.inv_step * P.pen_xs;
        EwIter I;
        I.a1 = 1.f + lam; I.a2 = -lam; I.cur = it & 1;
        I.br1 = d1 > thr1; I.br2 = d2 > thr2;
        I.sc1 = I.br1 ? 1.f - thr1 / d1 : 0.f; I.sc2 = I.br2 ? 1.f - thr2 / d2 : 0.f;
        ew_role(P, K.R, K.ES, I, bx, bp);
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
    // Grid barrier that ends phase S.  The stream ring is idle now: the shared sweep matrices are pulled over it while the
    // barrier is pending -- after the arrive, because its fence would otherwise also wait for those bulk copies.
    cbar();
    if (tid == 0) {
        P.cta_ns[2 * blockIdx.x] += globaltimer() - t_in;   // load balance of phase S (rn_cta_times)
        __threadfence();
        atomicAdd(P.bar, 1u);
        if (cta_sweeps_backward(P)) issue_matrix_loads(P);
    }
    dstamp(P, 0);
    if (tid == 0) {
        const unsigned int target = bar_count(P, it, 0);
        while (ld_acquire_u32(P.bar) < target) {}   // acquire + the CTA barrier below order the other threads' reads
    }
    cbar();
    dstamp(P, 1);
    // merged by one warp of a CTA that has no chain to sweep (or the fewest): it is off the critical path there
    const int pcta = P.K < (int)gridDim.x ? P.K : (int)gridDim.x - 1;
    if (it > 0 && (int)blockIdx.x == pcta && warp == 0) pinf_merge(P, it);
}

// Crown work is dealt from the last CTA down (those have the fewest chains, none when K < grid).  Item q goes to the CTA with
// slot = q mod grid.  First the bottom-crown nodes this rank owns (first pass: their S rows are what every rank's crown waits
// for, so they must not queue behind a chain's whole backward sweep), then the tiles of the second pass in id order (the upper
// crown first).  What overflows onto CTAs that sweep a chain are the last second-pass tiles, which wait for the S rows anyway.
struct CrownDeal { int n2, tile_w, n_tiles2, n_items; };
__device__ __forceinline__ CrownDeal crown_deal(const PArgs &P) {
    CrownDeal C;
    C.n2 = P.n_crown - P.n_owned;
    const int grid = (int)gridDim.x;
    C.tile_w = min(kTP / 2, max(1, (C.n2 + grid - 1) / grid));
    // (wider tiles that would keep the second pass off the CTAs with chains were measured slower at 8 GPUs on C3 -- the nodes of
    // a tile are summed one after the other: 8 967 against 9 462 iter/s with two nodes per tile -- so the overflow queues
    // behind a chain's backward sweep, which ends about when the S rows arrive anyway)
    C.n_tiles2 = (C.n2 + C.tile_w - 1) / C.tile_w;
    C.n_items = P.n_owned + C.n_tiles2;
    return C;
}
// does this CTA sweep anything backward / forward?  CTAs that do not leave the shared matrices where they are: no bulk copy
// without a waiter, less L2 traffic
__device__ __noinline__ bool cta_sweeps_backward(const PArgs &P) {
    if ((int)blockIdx.x < P.K) return true;
    if (P.n_crown == 0) return false;
    return (int)gridDim.x - 1 - (int)blockIdx.x < crown_deal(P).n_items;
}
struct CrownTiles { int tile_w, n_tiles; };
__device__ __forceinline__ CrownTiles crown_tiles(const PArgs &P) {
    // forward: as narrow as the CTAs WITHOUT a chain allow (a crown tile costs about as much as a chain's forward sweep, and up
    // to four columns of a tile cost the same as one: a tile on a chain's CTA would double that CTA's phase F); as narrow as the
    // grid allows when every CTA has chains
    CrownTiles C;
    const int grid = (int)gridDim.x, free_ctas = grid - min(P.K, grid);
    const int ctas = free_ctas > 0 ? free_ctas : grid;
    C.tile_w = min(kTP / 2, max(1, (P.n_crown + ctas - 1) / ctas));
    C.n_tiles = (P.n_crown + C.tile_w - 1) / C.tile_w;
    return C;
}
__device__ __forceinline__ bool cta_sweeps_forward(const PArgs &P) {
    return (int)blockIdx.x < P.K || (P.n_crown > 0 && (int)gridDim.x - 1 - (int)blockIdx.x < crown_tiles(P).n_tiles);
}

// second half of iteration `it`: the sweeps (phases B, C, F), this CTA's share of the prox distances, closing barrier.
// Three functions, each with little state of its own, so that the sweep steps below them keep their register budget.
__device__ __noinline__ void iter_backward(const PArgs &P, KState &K, int it) {
    const uint32_t mpar = (uint32_t)(it & 1);
    const int grid = (int)gridDim.x, j0 = (int)blockIdx.x, nK = P.K;
    // ---- phase B: backward sweep of the chains (operand blocks by bulk TMA, one chain ahead)
    dstamp(P, 2);
    if (threadIdx.x == 0 && j0 < nK) issue_chain_backward_loads(P, j0);
    if (threadIdx.x == 32 && !P.sh_mode && it + 1 < P.iters) l2_prefetch_next(P, K.R);
#pragma unroll 1
    for (int j = j0; j < nK; j += grid) chain_backward(P, j, j + grid < nK ? j + grid : -1, mpar, K.SP);
    dstamp(P, 10);
    // ---- phase C: backward sweep of the crown (tiles are dealt from the last CTA down: those have the fewest chains)
    if (P.n_crown > 0) {
        // the crown needs the heads of every chain: with several GPUs their q, r were stored into every rank's
        // table (chain_rscan), and this barrier spans the GPUs.  On one GPU only the CTAs that own crown
        // tiles wait, and only for the heads (a counter the chains bump after their r-scan): the crown overlaps the
        // rest of the chains' backward sweep
        const CrownDeal C = crown_deal(P);
        const int slot = grid - 1 - j0;
        // Pass 1 -- the bottom-crown nodes whose chains live on this rank: S_p (the sum over p's chain heads; awaits only p's
        // chains) goes to every rank's table, then p's own backward step follows at once.
#pragma unroll 1
        for (int q = slot; q < P.n_owned; q += grid) {
            const int p = P.own_lo + q;
            parent_sum(P, p, it);
            crown_backward(P, p, 1, mpar, K.SP, false, 0u, false);
        }
        // Pass 2 -- everything else of the crown needs every S row, this rank's and the peers': the CTAs that own tiles wait for
        // the counter parent_sum bumps on every rank; nobody else waits, and the crown overlaps the rest of the chains' sweep
        const unsigned int wait_s = P.s_base + (unsigned)P.n_bottom_active * (unsigned)(it + 1);   // counter value once every S row is in
        dstamp(P, 20);
        bool do_wait = true;
#pragma unroll 1
        for (int q = slot; q < C.n_items; q += grid) {
            if (q < P.n_owned) continue;
            const int tl = q - P.n_owned;
            crown_backward(P, tl * C.tile_w, min(C.tile_w, C.n2 - tl * C.tile_w), mpar, K.SP, do_wait, wait_s, true);
            do_wait = false;   // awaited once
        }
        dstamp(P, 21);
    }
    // G has done its work for this iteration: B takes its place (every step above ended with a CTA barrier).  With a
    // crown: grid barrier before phase F (the chains need the crown's L v), B's bulk copy between arrive and poll
    if (P.n_crown > 0) {
        cbar();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(P.bar, 1u);
            if (cta_sweeps_forward(P)) issue_b_load(P);
            if (j0 < nK) issue_chain_forward_loads(P, j0, 1);   // the staging area is free: this CTA's sweeps are done
            const unsigned int target = bar_count(P, it, 1);
            while (ld_acquire_u32(P.bar) < target) {}
        }
        cbar();
        dstamp(P, 22);
    } else if (threadIdx.x == 0 && cta_sweeps_forward(P)) issue_b_load(P);
}

// ---- phase F: forward sweep + prox boxes; leaves this CTA's sums of squared distances in K.s1, K.s2
__device__ __noinline__ void iter_forward(const PArgs &P, KState &K, int it) {
    const uint32_t mpar = (uint32_t)(it & 1);
    const int grid = (int)gridDim.x, j0 = (int)blockIdx.x, nK = P.K;
    const float *wxi = P.Wxi[it & 1], *wpsi = P.Wpsi[it & 1];
    if (threadIdx.x == 0 && j0 < nK) issue_chain_forward_loads(P, j0, P.n_crown > 0 ? 2 : 3);   // (with a crown: the rest went out at the barrier)
    K.s1 = 0; K.s2 = 0;
    if (P.n_crown > 0) {
        double c1 = 0, c2 = 0;
        const CrownTiles C = crown_tiles(P);
        const int n_crown = P.n_crown;
#pragma unroll 1
        for (int tl = grid - 1 - j0; tl < C.n_tiles; tl += grid)
            crown_forward(P, tl * C.tile_w, min(C.tile_w, n_crown - tl * C.tile_w), mpar, wxi, wpsi, c1, c2);
        if (P.rank == 0) { K.s1 = c1; K.s2 = c2; }   // the crown is replicated: it counts once in the global distances
    }
#pragma unroll 1
    for (int j = j0; j < nK; j += grid) chain_forward(P, j, j + grid < nK ? j + grid : -1, mpar, K.SP, wxi, wpsi, K.s1, K.s2);
    dstamp(P, 23);
}

__device__ __noinline__ void iter_close(const PArgs &P, KState &K, int it) {
    double *dsh = reinterpret_cast<double *>(smem_f(kOffDsh));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // this CTA's share of the two squared distances
        double s1 = K.s1, s2 = K.s2;
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) { dsh[warp] = s1; dsh[kPC / 32 + warp] = s2; }
        cbar();
        if (tid == 0) {
            double t1 = 0, t2 = 0;
            for (int w = 0; w < kPC / 32; w++) { t1 += dsh[w]; t2 += dsh[kPC / 32 + w]; }
            P.dist_part[2 * blockIdx.x] = t1; P.dist_part[2 * blockIdx.x + 1] = t2;
        }
    }
    const int last_pos = P.n_crown > 0 ? 2 : 1;
    const unsigned int target = bar_count(P, it, last_pos);
    dstamp(P, 28);
    if (P.n_ranks > 1) {
        // the one barrier across the GPUs per iteration (it also carries the prox distances).  Arrive first -- the fence would
        // otherwise also wait for the bulk copies -- then refill the stream ring for iteration it+1 (the sweep region is dead:
        // every warp passed the CTA barrier inside the arrive), then poll the peers' flags
        cross_arrive(P, target, P.epoch0 + (unsigned)it + 1u, it);
        if (warp == kLoaderWarp && it + 1 < P.iters) {
            if (!P.sh_mode) loader_role(P, K.R, K.LS, it + 1, 1);
            else if (lane == 0) sh_issue_matrices(P);
        }
        cross_wait(P, P.epoch0 + (unsigned)it + 1u);
    } else {
        // closing grid barrier, run by the loader warp: arrive first (the fence would otherwise also wait for the bulk
        // copies), then refill the stream ring for iteration it+1 -- the sweep region is dead, the factor matrices are
        // constants -- and only then poll
        cbar();
        if (warp == kLoaderWarp) {
            if (lane == 0) { __threadfence(); atomicAdd(P.bar, 1u); }
            __syncwarp();
            if (it + 1 < P.iters) {
                if (!P.sh_mode) loader_role(P, K.R, K.LS, it + 1, 1);
                else if (lane == 0) sh_issue_matrices(P);   // shared-factor mode: G, L' of the next iteration
                __syncwarp();
            }
            if (lane == 0) while (ld_acquire_u32(P.bar) < target) {}   // acquire + the CTA barrier below order the other threads' reads
            __syncwarp();
            // the global distances of this iteration's prox (cublasSnrm2, :792, :810) for the next iteration's element-wise
            // pass: every CTA adds the per-CTA partial sums in the same fixed order
            double p1 = 0, p2 = 0;
            for (int k = lane; k < (int)gridDim.x; k += 32) { p1 += __ldcg(P.dist_part + 2 * k); p2 += __ldcg(P.dist_part + 2 * k + 1); }
            for (int o = 16; o > 0; o >>= 1) { p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o); }
            if (lane == 0) { float *sd = smem_f(kOffSd); sd[0] = (float)sqrt(p1); sd[1] = (float)sqrt(p2); }
        }
        cbar();
    }
    dstamp(P, 29);
}

__global__ void __launch_bounds__(kPT, 1) k_apg_persistent(const __grid_constant__ PArgs P) {
    // The non-inlined functions get the arguments through a reference; reads of the kernel-parameter space behind a
    // generic pointer cost about a microsecond each (ncu: 12 % of all stall samples sat on them), so they are handed a
    // shared-memory copy instead.
    static_assert(sizeof(PArgs) <= kPArgsWords * 4 && sizeof(PArgs) % 4 == 0, "PArgs copy");
    {
        const int *src = reinterpret_cast<const int *>(&P);
        int *dst = reinterpret_cast<int *>(smem_f(kOffPArgs));
        for (int k = threadIdx.x; k < (int)(sizeof(PArgs) / 4); k += kPC) dst[k] = src[k];
    }
    const PArgs &PS = *reinterpret_cast<const PArgs *>(smem_f(kOffPArgs));
    const int tid = threadIdx.x;
    if (tid == 0) {
        const Pipe M = pipe_smem();
        for (int s = 0; s < kPStages; s++) { mbar_init(&M.full[s], 1); mbar_init(&M.empty[s], kGemvWarps); }
        for (int s = 0; s < kVecSlots; s++) { mbar_init(&M.vfull[s], 32); mbar_init(&M.vempty[s], kEwWarps); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&M.wfull[s], kEwWarps); mbar_init(&M.wempty[s], kGemvWarps);
            mbar_init(&M.rfull[s], kGemvWarps); mbar_init(&M.rempty[s], kEwWarps);
        }
        for (int s = 0; s < 8; s++) mbar_init(&M.rempty[2 + s], 1);   // = SweepSmem::mfull, sfull
        mbar_fence_init();
        unsigned long long *c = clk_smem();
        for (int k = 0; k < 32; k++) c[k] = 0;
        c[32] = globaltimer(); c[33] = blockIdx.x == P.clock_cta ? 1 : 0;
        reinterpret_cast<int *>(smem_f(kOffMeta))[0] = -1;
        smem_f(kOffSd)[0] = 0.f; smem_f(kOffSd)[1] = 0.f;   // no prox before iteration 0
    }
    {   // u_prev, uhat_prev, x_cur: read by every forward scan, fixed for the launch
        float *fix = smem_f(kOffFix);
        for (int k = tid; k < P.nu; k += kPC) { fix[k] = __ldg(P.uprev + k); fix[kVStride + k] = __ldg(P.uhat_prev + k); }
        for (int k = tid; k < P.nx; k += kPC) fix[2 * kVStride + k] = __ldg(P.xcur + k);
    }
    __syncthreads();

    KState K;
    // this CTA's slice of the stream: matrix units [u_begin, u_end) in node-major order
    const long long n_units = (long long)P.nodes * P.n_mats;
    K.R.u_begin = (int)(n_units * blockIdx.x / gridDim.x);
    K.R.u_end = (int)(n_units * (blockIdx.x + 1) / gridDim.x);
    K.R.node_first = K.R.u_begin < K.R.u_end ? K.R.u_begin / P.n_mats : 0;
    K.R.node_last = K.R.u_begin < K.R.u_end ? (K.R.u_end - 1) / P.n_mats : -1;
    K.SP = StagePhase{0u, 0u, 0u, 0u};
    K.GS = GemvState{0, 0u, 0, 0u, 0, 0u};
    K.ES = EwState{0, 0u, 0, 0u, 0, 0u};
    K.LS = LoaderState{0, 0u, 0, 0u, 0};
    K.SH = ShState{0u};
    K.s1 = 0; K.s2 = 0;

    const int iters = P.iters;
    if (P.sh_mode && tid == 0 && iters > 0) sh_issue_matrices(PS);   // later iterations: at the closing barrier of the one before
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        iter_stream(PS, K, it);
        iter_backward(PS, K, it);
        iter_forward(PS, K, it);
        iter_close(PS, K, it);
    }
    if (P.n_ranks > 1 && blockIdx.x == 0 && tid == 0) {   // k_finalize reads the global sums from slot 0
        double t1 = 0, t2 = 0;
        const double *ds = P.dslot_peer[P.rank] + 2 * ((P.iters + 1) & 1);
        for (int r = 0; r < P.n_ranks; r++) { t1 += __ldcg(ds + 4 * r); t2 += __ldcg(ds + 4 * r + 1); }
        P.dist_part[0] = t1; P.dist_part[1] = t2;
    }
    if (blockIdx.x == P.clock_cta && tid == 0) {
        const unsigned long long *c = clk_smem();
        for (int k = 0; k < 32; k++) P.phase_ns[k] += c[k];
    }
    if (tid == 0) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); P.cta_ns[2 * blockIdx.x + 1] = smid; }
    if (blockIdx.x == 0 && tid == 0) *P.iter_dev = P.iters - 1;   // k_finalize finishes iteration iters-1
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize of k_apg_persistent: one value per device for the whole process
static rn_status raise_persist_smem_limit(Handle *h, int bytes);
struct SweepLayout {
    int oG, oOm, oL, oX1, oY, oV, oScr2, oStg, oXb, end;   // shared-memory float offsets
    int pG, pOm, pL, pB, pLt, pack_floats, sB, sLt;   // float offsets inside the pack (G | OmegaBar | L | B | L'), sizes of B and L'
    int nxp, nup, nvp;
};
static int pad4(long long n) { return (int)((n + 3) & ~3LL); }
static SweepLayout sweep_layout(const Handle *h) {
    const int nx = h->d.nx, nu = h->d.nu, nv = h->d.nv, T = h->d.N - h->chain_stage;
    SweepLayout Y{};
    Y.nxp = pad4(nx); Y.nup = pad4(nu); Y.nvp = pad4(nv);
    const int sG = pad4((long long)nv * nx), sOm = pad4((long long)nv * nv), sL = pad4((long long)nu * nv), sB = pad4((long long)nx * nu);
    Y.pG = 0; Y.pOm = sG; Y.pL = sG + sOm; Y.pB = sG + sOm + sL; Y.pLt = Y.pB + sB; Y.pack_floats = Y.pLt + sL;
    Y.sB = sB; Y.sLt = sL;
    Y.oG = kOffSweep; Y.oOm = Y.oG + std::max(sG, sB); Y.oL = Y.oOm + sOm;     // B overlays G
    Y.oX1 = Y.oL + sL;
    Y.oY = Y.oX1 + std::max(nv + nx, nu) * kTP;
    Y.oV = Y.oY + std::max(std::max(nv, nu), nx) * kTP;
    Y.oScr2 = Y.oV + std::max(nv, nu) * kTP;
    Y.oStg = Y.oScr2 + 128 * kTP;
    const int cs = h->chain_stage;
    Y.oXb = Y.oStg + std::max(5 * T * Y.nvp, 2 * T * Y.nup + T * Y.nxp + cs * (2 * Y.nup + Y.nxp));
    Y.end = Y.oXb + nx * kTP;
    return Y;
}
// matrix ring of phase S: whole factor matrices per stage when three of them fit next to the sweep region's size,
// otherwise as many columns as a stage holds (stage payload >= 16 KB, at least 3 stages)
struct RingLayout { int n_stages, stage_stride, cols_per_chunk, end; };
static RingLayout ring_layout(const Handle *h) {
    const int nx = h->d.nx, nu = h->d.nu, nv = h->d.nv;
    const int budget = std::max(sweep_layout(h).end, kOffRing + 6 * (kPStageMinFloats + 32)) - kOffRing;   // floats for the ring
    const int whole = ((nv * std::max(2 * nx, nu) + 31) & ~31) + 32;
    RingLayout R{};
    if (3 * whole <= budget) {
        R.stage_stride = whole;
        R.n_stages = std::min(kPStages, budget / whole);
        R.cols_per_chunk = std::max(2 * nx, nu);
    } else {
        const int payload = std::max(kPStageMinFloats, ((budget / 6 - 32) / 32) * 32);
        R.stage_stride = payload + 32;
        R.n_stages = std::min(kPStages, std::max(3, budget / R.stage_stride));
        R.cols_per_chunk = payload / nv;
    }
    R.end = kOffRing + R.n_stages * R.stage_stride;
    return R;
}
// RN_FACTORS_SHARED, phase S: G | L' | c columns [nx][kTP] | g columns [nu][kTP] | products [nv][kTP] | split-K scratch |
// vector buffer of one tile of `tile` nodes (Hx, w, z, y of both dual blocks + the diagonals, each with 8 floats of slack
// for the 16-byte window of its bulk copy).  It overlays the sweep region like the stream ring does.
struct SharedLayout { int oLt, oC, oGc, oY, oScr2, oVec, tile, end; };
static int shared_vec_floats(const rn_dims &d, int tile) {
    return 4 * (pad4((long long)tile * 2 * d.nx) + 8) + 4 * (pad4((long long)tile * d.nu) + 8) + pad4((long long)tile * (2 * d.nx + d.nu)) + 8;
}
static SharedLayout shared_layout(const Handle *h) {
    const rn_dims &d = h->d;
    SharedLayout S{};
    S.oLt = kOffSweep + ((pad4((long long)d.nv * d.nx) + 31) & ~31);
    S.oC = S.oLt + ((pad4((long long)d.nu * d.nv) + 31) & ~31);
    S.oGc = S.oC + d.nx * kTP; S.oY = S.oGc + d.nu * kTP; S.oScr2 = S.oY + d.nv * kTP;
    S.oVec = (S.oScr2 + 128 * kTP + 31) & ~31;
    const int cap = 227 * 1024 / 4 - 32;
    S.tile = kTP;
    while (S.tile > 1 && S.oVec + shared_vec_floats(d, S.tile) > cap) S.tile--;
    S.end = S.oVec + shared_vec_floats(d, S.tile);
    return S;
}
static size_t persist_smem_bytes(const Handle *h) {
    int end = std::max(ring_layout(h).end, sweep_layout(h).end);
    if (h->factor_mode == RN_FACTORS_SHARED) end = std::max(end, shared_layout(h).end);
    return (size_t)end * 4 + 128;
}

bool persistent_supported(const Handle *h) {
    const rn_dims &d = h->d;
    if (std::max(2 * d.nx, std::max(d.nu, d.nv)) > kDimMax) return false;
    const int cs = h->chain_stage;
    if (cs >= d.N || cs > kMaxCs) return false;                       // needs a non-branching tail; bounded crown depth
    if (h->h_nps[cs] != d.K) return false;                             // the tail's chains are the scenarios
    if (d.N - cs + (cs > 0 ? 1 : 0) > kTP) return false;               // a chain (+ its parent's column) is one tile
    if (ring_layout(h).cols_per_chunk < 1) return false;
    return persist_smem_bytes(h) <= 227 * 1024;
}

// dst[pos(node)][dimp] = src[node][dim]: chain-major copies of the per-solve vectors (beta, uhat, e)
__global__ void k_to_chain_major(int nodes, int dim, int dimp, const int *__restrict__ pos, const float *__restrict__ src,
                                 float *__restrict__ dst) {
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const size_t row = (size_t)pos[i] * dimp;
    for (int t = threadIdx.x; t < dim; t += blockDim.x) dst[row + t] = src[(size_t)i * dim + t];
}

// Exchange buffer of this rank (one cudaMalloc, so that one CUDA IPC handle maps it into the peer processes):
//   S_g table [n_crown*nvp] | S_r table [n_crown*nvp] (rows padded to 16 bytes) | beta rows of the crown [n_crown*nv] | distance slots [kMaxRanks][2][2] doubles | flags [kMaxRanks] | S-row counter | err
struct XchgLayout { size_t sq, sr, beta, dslot, flags, sctr, err, bytes; };
static XchgLayout xchg_layout(const Handle *h) {
    XchgLayout X{};
    const size_t nc = (size_t)std::max(h->h_cum[h->chain_stage], 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~size_t(255); return at; };
    X.sq = take(nc * pad4(h->d.nv) * 4 + 64); X.sr = take(nc * pad4(h->d.nv) * 4 + 64); X.beta = take(nc * h->d.nv * 4); X.dslot = take(kMaxRanks * 4 * 8);
    X.flags = take(kMaxRanks * 4); X.sctr = take(4); X.err = take(4); X.bytes = o;
    return X;
}
size_t xchg_err_offset(const Handle *h) { return xchg_layout(h).err; }
// The one per-SOLVE quantity that couples the ranks: zeta (hence beta) of a bottom-crown node sums over its children
// (calculateZeta, Utilities.cu:100-131), and those live on one rank only -- the partition is aligned to these nodes.  The
// owner's beta row is exact; it is pushed into every rank's staging table (push), and after a barrier across the ranks every
// rank copies the rows it does not own into its beta (pull).
__global__ void k_push_beta(int bottom0, int n_bottom, int nv, int n_ranks, const int *__restrict__ child_count,
                            const float *__restrict__ beta, float *s0, float *s1, float *s2, float *s3, float *s4, float *s5,
                            float *s6, float *s7) {
    const int p = bottom0 + blockIdx.x;
    if ((int)blockIdx.x >= n_bottom || child_count[p] == 0) return;
    float *dst[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
    for (int t = threadIdx.x; t < nv; t += blockDim.x) {
        const float v = beta[(size_t)p * nv + t];
        for (int r = 0; r < n_ranks; r++) dst[r][(size_t)p * nv + t] = v;
    }
}
__global__ void k_pull_beta(int bottom0, int n_bottom, int nv, const int *__restrict__ child_count, const float *__restrict__ stage,
                            float *__restrict__ beta) {
    const int p = bottom0 + blockIdx.x;
    if ((int)blockIdx.x >= n_bottom || child_count[p] != 0) return;
    for (int t = threadIdx.x; t < nv; t += blockDim.x) beta[(size_t)p * nv + t] = stage[(size_t)p * nv + t];
}
rn_status dist_crown_beta(Handle *h, int pull) {
    const int cs = h->chain_stage;
    if (h->dist_world <= 1 || cs <= 0) return RN_OK;
    const XchgLayout X = xchg_layout(h);
    const int bottom0 = h->h_cum[cs - 1], n_bottom = h->h_cum[cs] - bottom0;
    if (pull) {
        k_pull_beta<<<n_bottom, 128, 0, h->stream>>>(bottom0, n_bottom, h->d.nv, h->t.child_count,
                                                     reinterpret_cast<const float *>(static_cast<char *>(h->xchg) + X.beta), h->beta);
    } else {
        float *dst[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int r = 0; r < h->dist_world; r++) {
            if (!h->xchg_peer[r]) return fail(h, RN_ERR_STATE, "rank %d's exchange buffer is not connected (rn_dist_connect)", r);
            dst[r] = reinterpret_cast<float *>(static_cast<char *>(h->xchg_peer[r]) + X.beta);
        }
        k_push_beta<<<n_bottom, 128, 0, h->stream>>>(bottom0, n_bottom, h->d.nv, h->dist_world, h->t.child_count, h->beta, dst[0], dst[1],
                                                     dst[2], dst[3], dst[4], dst[5], dst[6], dst[7]);
    }
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

rn_status ensure_xchg(Handle *h) {
    if (h->xchg) return RN_OK;
    const XchgLayout X = xchg_layout(h);
    RN_CUDA(h, cudaMalloc(&h->xchg, X.bytes));
    h->allocs.push_back(h->xchg);
    h->device_bytes += X.bytes;
    RN_CUDA(h, cudaMemset(h->xchg, 0, X.bytes));
    for (int r = 0; r < kMaxRanks; r++) h->xchg_peer[r] = nullptr;
    h->xchg_peer[h->dist_rank] = h->xchg;
    return RN_OK;
}

// The shared matrices of the sweeps are outputs of the factor step (G = Bbar', OmegaBar, L and L' change with the
// null-space basis, rn_set_null_space): the pack is re-copied on `st` whenever a factor step has run since the last launch.
static rn_status refresh_pack(Handle *h, cudaStream_t st) {
    const rn_dims &d = h->d;
    const SweepLayout Y = sweep_layout(h);
    const size_t f = sizeof(float);
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pG, h->G, (size_t)d.nv * d.nx * f, cudaMemcpyDeviceToDevice, st));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pOm, h->OmegaBar, (size_t)d.nv * d.nv * f, cudaMemcpyDeviceToDevice, st));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pL, h->L, (size_t)d.nu * d.nv * f, cudaMemcpyDeviceToDevice, st));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pB, h->B, (size_t)d.nx * d.nu * f, cudaMemcpyDeviceToDevice, st));
    RN_CUDA(h, cudaMemcpyAsync(h->sweep_pack + Y.pLt, h->Lt, (size_t)d.nv * d.nu * f, cudaMemcpyDeviceToDevice, st));
    h->pack_dirty = false;
    return RN_OK;
}

rn_status persistent_prepare(Handle *h) {
    if (h->persist_ready) return RN_OK;
    const rn_dims &d = h->d;
    const size_t n = d.nodes;
    const int cs = h->chain_stage, n_crown = h->h_cum[cs], T = d.N - cs;
    const SweepLayout Y = sweep_layout(h);
    for (int k = 0; k < 4; k++) RN_CHECK(dev_alloc(h, &h->part[k], n * Y.nvp));
    RN_CHECK(dev_alloc(h, &h->cm_c, n * Y.nxp)); RN_CHECK(dev_alloc(h, &h->cm_lv, n * Y.nup));
    RN_CHECK(dev_alloc(h, &h->cm_beta, n * Y.nvp)); RN_CHECK(dev_alloc(h, &h->cm_uhat, n * Y.nup)); RN_CHECK(dev_alloc(h, &h->cm_e, n * Y.nxp));
    RN_CHECK(ensure_xchg(h));
    // [0] grid barrier, [32] published-S counter (its own 128-byte line), [64 ...] per bottom-crown node: heads in
    RN_CHECK(dev_alloc(h, &h->grid_bar, 64 + (size_t)std::max(n_crown, 1)));
    RN_CHECK(dev_alloc(h, &h->head_q, (size_t)d.K * d.nv)); RN_CHECK(dev_alloc(h, &h->head_r, (size_t)d.K * d.nv));
    RN_CHECK(dev_alloc(h, &h->phase_ns, 32));
    RN_CHECK(dev_alloc(h, &h->cta_ns, 2 * 1024));
    // G | OmegaBar | L | B | L', each padded to 16 bytes: the bulk copies of the sweeps (and of phase S in shared-factor mode)
    RN_CHECK(dev_alloc(h, &h->sweep_pack, (size_t)Y.pack_floats));
    h->pack_dirty = true;   // filled by refresh_pack at the next launch (and again after every factor step)
    // chain-major row of every node: crown nodes keep their id, chain j / stage cs+s sits at n_crown + j T + s
    std::vector<int> pos(n);
    for (int i = 0; i < d.nodes; i++) {
        const int s = h->h_stages[i];
        pos[i] = s < cs ? i : n_crown + (i - h->h_cum[s]) * T + (s - cs);
    }
    RN_CHECK(dev_alloc(h, &h->pos_dev, n));
    RN_CUDA(h, cudaMemcpyAsync(h->pos_dev, pos.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    // descendant id range of every crown node at every later stage up to the chain heads (children are contiguous)
    // At stage cs the range is in chain indices; with several GPUs the host supplies the GLOBAL chain ranges there.
    std::vector<int> rng((size_t)std::max(n_crown, 1) * (kMaxCs + 1) * 2, 0);
    for (int i = 0; i < n_crown; i++) {
        int lo = i, hi = i + 1;
        for (int s = h->h_stages[i] + 1; s <= cs; s++) {
            const int nlo = h->h_child_first[lo], nhi = h->h_child_first[hi - 1] + h->h_child_count[hi - 1];
            lo = nlo; hi = nhi;
            // (stage cs is not tabulated: the heads enter through the S rows of the bottom-crown nodes, parent_sum)
            if (s == cs) break;
            rng[((size_t)i * (kMaxCs + 1) + s) * 2] = lo; rng[((size_t)i * (kMaxCs + 1) + s) * 2 + 1] = hi;
        }
    }
    // path root -> node of every crown node
    std::vector<int> cpath((size_t)std::max(n_crown, 1) * kMaxCs, 0);
    for (int i = 0; i < n_crown; i++) {
        int a = i;
        for (int k = h->h_stages[i]; k >= 0 && a >= 0; k--) { if (k < kMaxCs) cpath[(size_t)i * kMaxCs + k] = a; a = h->h_parent[a]; }
    }
    RN_CHECK(dev_alloc(h, &h->crown_path, cpath.size()));
    RN_CUDA(h, cudaMemcpyAsync(h->crown_path, cpath.data(), cpath.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    RN_CHECK(dev_alloc(h, &h->crown_rng, rng.size()));
    RN_CUDA(h, cudaMemcpyAsync(h->crown_rng, rng.data(), rng.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    size_t smem = persist_smem_bytes(h);
    if ((size_t)shared_layout(h).end * 4 + 128 <= 227 * 1024) smem = std::max(smem, (size_t)shared_layout(h).end * 4 + 128);
    RN_CHECK(raise_persist_smem_limit(h, (int)smem));   // for the occupancy query and every launch of this handle
    int per_sm = 0;
    RN_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_apg_persistent, kPT, smem));
    if (per_sm < 1) return fail(h, RN_ERR_INVALID, "persistent kernel does not fit on an SM (%zu B shared memory)", smem);
    int coop = 0;
    RN_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) return fail(h, RN_ERR_CUDA, "device does not support cooperative launch");
    h->persist_ready = true;
    return RN_OK;
}

// iterations 0 .. iters-1 up to (and excluding) the last finalisation; the caller runs k_finalize afterwards
rn_status persistent_launch(Handle *h, cudaStream_t st, int iters) {
    const rn_dims &d = h->d;
    h->persist_grid = std::min(h->sm_count, std::min(h->dist_slots, h->pinf_slots));
    if (h->grid_limit > 0) h->persist_grid = std::min(h->persist_grid, h->grid_limit);   // rn_set_grid_limit
    PArgs P{};
    P.parent = h->t.parent; P.child_first = h->t.child_first; P.child_count = h->t.child_count; P.omega_idx = h->t.omega_idx;
    P.cum = h->cum_dev; P.stages = h->t.stages; P.crown_rng = h->crown_rng; P.pos = h->pos_dev; P.crown_path = h->crown_path;
    // the reference's branching stages (:699-719) are those of the WHOLE tree: a rank of a partition holds K of K_glob chains
    // below the replicated crown, and a count of its own stages would call the chain heads' stage non-branching as soon as
    // K < nodes of the last crown stage (eight GPUs on C3: 60 < 80) -- same mathematics, different order of the additions
    P.branch_mask = 0;
    for (int st_ = 1; st_ < d.N && st_ < 32; st_++) {
        auto n_at = [&](int s) { return (h->dist_world > 1 && s >= h->chain_stage) ? h->dist_K_glob : h->h_cum[s + 1] - h->h_cum[s]; };
        if (n_at(st_) > n_at(st_ - 1)) P.branch_mask |= 1u << st_;
    }
    P.N = d.N; P.cs = h->chain_stage; P.K = d.K; P.nodes = d.nodes; P.n_crown = h->h_cum[h->chain_stage];
    P.n_mats = h->factor_mode == RN_FACTORS_FULL ? 4 : 2;
    P.df_mode = h->factor_mode != RN_FACTORS_FULL ? 1 : 0;           // v = -1/2 Omega r
    P.sh_mode = h->factor_mode == RN_FACTORS_SHARED ? 1 : 0;          // D xi = G c, F psi = L' g: no per-node matrix
    {
        // share of the streamed matrices kept L2-resident across iterations: RN_L2_KEEP x L2 size (x this handle's share of
        // the SMs, rn_set_grid_limit: lanes share the L2) over the bytes one iteration streams
        static const char *env = getenv("RN_L2_KEEP");
        static const double keep_l2 = env ? atof(env) : kL2KeepDefault;
        const double bytes = h->factor_mode == RN_FACTORS_FULL ? (double)h->factor_full_bytes : (double)h->factor_df_bytes;
        P.l2_keep = 0.f;
        if (keep_l2 > 0 && !P.sh_mode && bytes > 0)
            P.l2_keep = (float)std::min(1.0, keep_l2 * (double)h->l2_bytes * h->persist_grid / h->sm_count / bytes);
        static const char *penv = getenv("RN_L2_PREFETCH");
        static const double pref_l2 = penv ? atof(penv) : kL2PrefDefault;
        P.l2_pref = 0.f;
        if (pref_l2 > 0 && !P.sh_mode && bytes > 0)
            P.l2_pref = (float)std::min(1.0 - P.l2_keep, pref_l2 * (double)h->l2_bytes * h->persist_grid / h->sm_count / bytes);
    }
    {
        const SharedLayout SL = shared_layout(h);
        P.sh_tile = SL.tile; P.sh_oLt = SL.oLt; P.sh_oC = SL.oC; P.sh_oGc = SL.oGc; P.sh_oY = SL.oY; P.sh_oScr2 = SL.oScr2; P.sh_oVec = SL.oVec;
    }
    P.iters = iters; P.nx = d.nx; P.nu = d.nu; P.nv = d.nv;
    const RingLayout RL = ring_layout(h);
    P.cols_per_chunk = RL.cols_per_chunk; P.n_stages = RL.n_stages; P.stage_stride = RL.stage_stride;
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
    P.V = h->V; P.U = h->U; P.X = h->X;
    {
        const XchgLayout X = xchg_layout(h);
        P.n_ranks = std::max(h->dist_world, 1); P.rank = h->dist_world > 1 ? h->dist_rank : 0;
        P.K_glob = h->dist_world > 1 ? h->dist_K_glob : d.K; P.chain_off = h->dist_world > 1 ? h->dist_chain_off : 0;
        for (int r = 0; r < P.n_ranks; r++) {
            char *base = static_cast<char *>(h->xchg_peer[r]);
            if (!base) return fail(h, RN_ERR_STATE, "rank %d's exchange buffer is not connected (rn_dist_connect)", r);
            P.Sg_peer[r] = reinterpret_cast<float *>(base + X.sq); P.Sr_peer[r] = reinterpret_cast<float *>(base + X.sr);
            P.dslot_peer[r] = reinterpret_cast<double *>(base + X.dslot);
            P.xflag_peer[r] = reinterpret_cast<unsigned int *>(base + X.flags);
            P.s_ctr_peer[r] = reinterpret_cast<unsigned int *>(base + X.sctr);
        }
        char *own = static_cast<char *>(h->xchg);
        P.xerr = reinterpret_cast<int *>(own + X.err);
        P.hg = h->head_q; P.hr = h->head_r;
        const int cs_ = h->chain_stage;
        P.bottom0 = cs_ > 0 ? h->h_cum[cs_ - 1] : 0; P.n_bottom = cs_ > 0 ? h->h_cum[cs_] - h->h_cum[cs_ - 1] : 0;
        P.n_owned = 0; P.own_lo = P.bottom0;
        for (int b = 0; b < P.n_bottom; b++)
            if (h->h_child_count[P.bottom0 + b] > 0) { if (P.n_owned == 0) P.own_lo = P.bottom0 + b; P.n_owned++; }
        // rows every rank's S table receives per iteration: every bottom-crown node that has chains somewhere (all of them in a
        // scenario tree: a node above the chain stage has children); the counters run on across launches, like the epochs
        P.n_bottom_active = P.n_ranks > 1 ? P.n_bottom : P.n_owned;
        P.s_base = h->s_total;
        h->s_total += (unsigned)P.n_bottom_active * (unsigned)iters;
        for (int b = 0; b < P.n_owned; b++)   // the owned bottom-crown nodes are one contiguous id range (contiguous chain ranges)
            if (h->h_child_count[P.own_lo + b] == 0) return fail(h, RN_ERR_INVALID, "the bottom-crown nodes with local chains are not contiguous");
        RN_CUDA(h, cudaMemsetAsync(own + X.err, 0, sizeof(int), st));   // a timed-out wait of an earlier launch does not poison this one
        P.epoch0 = h->xepoch;
        h->xepoch += (unsigned)iters + 2u;   // every rank runs the same launches: the epochs stay aligned
    }
    if (h->pinf4_cap < iters) {
        h->pinf4_cap = iters + 8;
        RN_CHECK(dev_alloc(h, &h->pinf4, (size_t)h->pinf4_cap * 4));
    }
    P.pinf4 = h->pinf4;
    P.cm_c = h->cm_c; P.cm_lv = h->cm_lv; P.cm_beta = h->cm_beta; P.cm_uhat = h->cm_uhat; P.cm_e = h->cm_e;
    P.dist_part = h->dist_part; P.pinf = h->pinf; P.pinf_part = h->pinf_part;
    P.lambda_tab = h->lambda_tab; P.bar = h->grid_bar; P.par_ctr = h->grid_bar + 64; P.iter_dev = h->iter_dev; P.phase_ns = h->phase_ns; P.cta_ns = h->cta_ns;
    P.step = h->step; P.inv_step = 1 / h->step; P.pen_x = h->pen_x; P.pen_xs = h->pen_xs;
    const SweepLayout Y = sweep_layout(h);
    P.oG = Y.oG; P.oOm = Y.oOm; P.oL = Y.oL; P.oX1 = Y.oX1; P.oY = Y.oY; P.oV = Y.oV; P.oScr2 = Y.oScr2; P.oStg = Y.oStg; P.oXb = Y.oXb;
    P.pG = Y.pG; P.pOm = Y.pOm; P.pL = Y.pL; P.pB = Y.pB; P.pLt = Y.pLt;
    P.bG = (unsigned)(Y.pOm - Y.pG) * 4u; P.bOm = (unsigned)(Y.pL - Y.pOm) * 4u; P.bL = (unsigned)(Y.pB - Y.pL) * 4u;
    P.bB = (unsigned)Y.sB * 4u; P.bLt = (unsigned)Y.sLt * 4u;
    P.nxp = Y.nxp; P.nup = Y.nup; P.nvp = Y.nvp;
    // chain-major copies of the per-solve vectors the sweeps stage with bulk copies
    k_to_chain_major<<<d.nodes, 128, 0, st>>>(d.nodes, d.nv, Y.nvp, h->pos_dev, h->beta, h->cm_beta);
    k_to_chain_major<<<d.nodes, 128, 0, st>>>(d.nodes, d.nu, Y.nup, h->pos_dev, h->uhat, h->cm_uhat);
    k_to_chain_major<<<d.nodes, 128, 0, st>>>(d.nodes, d.nx, Y.nxp, h->pos_dev, h->e, h->cm_e);
    h->launches += 3;
    RN_CUDA(h, cudaMemsetAsync(h->grid_bar, 0, (64 + (size_t)std::max(P.n_crown, 1)) * sizeof(unsigned int), st));
    RN_CUDA(h, cudaMemsetAsync(h->dist_part, 0, 2 * sizeof(double) * h->persist_grid, st));
    if (h->pack_dirty) RN_CHECK(refresh_pack(h, st));
    // the attribute belongs to the function and the device, not to the handle: a handle of a smaller problem prepared later
    // must not lower the limit under this one.  A running maximum per device, raised only when a launch needs more: setting the
    // attribute at every launch serialised the lanes of a closed-loop study (four handles solving side by side on one GPU ran at
    // the pace of one)
    RN_CHECK(raise_persist_smem_limit(h, (int)persist_smem_bytes(h)));
    void *args[] = {&P};
    RN_CUDA(h, cudaLaunchCooperativeKernel((const void *)k_apg_persistent, dim3(h->persist_grid), dim3(kPT), args,
                                           persist_smem_bytes(h), st));
    return RN_OK;
}

static rn_status raise_persist_smem_limit(Handle *h, int bytes) {
    static std::mutex mu;
    static int limit[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    const int dev = h->device >= 0 && h->device < 64 ? h->device : 0;
    if (bytes > limit[dev]) {
        RN_CUDA(h, cudaFuncSetAttribute(k_apg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        limit[dev] = bytes;
    }
    return RN_OK;
}

}  // namespace rn
