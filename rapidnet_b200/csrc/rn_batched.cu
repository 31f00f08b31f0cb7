// rn_batched.cu -- the tree sweeps of SmpcController::solveStep as GEMMs across ALL nodes (RN_SWEEP_BATCHED).
//
// Reference: /root/reference/src/SmpcController.cu:593-747 multiplies, stage by stage, every node's vectors with matrices
// that are the same for all nodes of the tree up to a scalar: G = Bbar', Omega_i = OmegaBar / p_i, Theta_i = -1/2 Omega_i Bbar',
// L, B (Engine.cu:707-747).  The stage recursion itself only adds vectors (q = c + q_child, sigma = beta + r_child,
// u = uhat + (u_par - uhat_par) + L v, x = x_par + e + B u), so the sweep factors into
//
//   scan  q_bar_i = sum_children q_c, q = c + q_bar                (k_bat_q_chain / k_bat_q_stage)
//   GEMM  Y1 = G      [nv x nx] * Q_bar [nx x nodes]                        (k_sgemm)
//   scan  sigma = beta + sum_children r_c, r = (sigma + a) + Y1, X = -1/2 (sigma + Y1)   [df: -1/2 r]
//   GEMM  Y2 = OmegaBar [nv x nv] * X   [nv x nodes];   v = Y2 / p + b      (Theta q_bar = -1/2 Omega (G q_bar))
//   GEMM  Y3 = L      [nu x nv] * V     [nv x nodes]
//   scan  u = ((uhat + u_par) - uhat_par) + Y3                      (the reference's two orders, :683-728)
//   GEMM  Y4 = B      [nx x nu] * U     [nu x nodes]
//   scan  x = (x_par + e) + Y4, then Hx, Hu, t = Hx + w / step, box projections, distance partial sums
//
// -- four GEMMs whose N is the number of nodes: the north-star's "stage nodes share a factor matrix, so the sweep becomes a
// true GEMM across nodes".  The scans run one thread per (chain, element) down the non-branching tail and one small
// launch per branching stage above it.  This is the path of problems the persistent kernel does not hold in shared memory
// (BASELINE config[4]: 4 x Barcelona, OmegaBar alone is 602 KB); every shared matrix is read once per tile from L2 instead
// of once per node.  The products are fp32 FFMA (3xTF32 tensor-core GEMMs would change the rounding of the iterates).
#include <algorithm>

#include "rn_internal.h"
#include "rn_device.cuh"

namespace rn {

// ---------------------------------------------------------------------------------------------------------------------
// C[M x N] = A[M x K] * B[K x N], all column-major without padding (ldA = M, ldB = K, ldC = M): A is a shared matrix of
// the Engine, the columns of B / C are the nodes' vectors in the reference's node-major layout.
// 64 x 64 tile per CTA, 16 k per step, 256 threads x (4 x 4) outputs; the next tile's operands are fetched into registers
// while the current one is multiplied.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBM = 64, kBN = 64, kBK = 16;
__global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, const float *__restrict__ A, const float *__restrict__ B,
                                               float *__restrict__ C) {
    __shared__ __align__(16) float As[kBK][kBM + 4];
    __shared__ __align__(16) float Bs[kBK][kBN + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    // loaders: A tile 64 x 16 = 4 per thread (rows contiguous), B tile 16 x 64 = 4 per thread (k contiguous)
    const int am = t & 63, ak = t >> 6;          // A: row am, columns ak, ak + 4, ak + 8, ak + 12
    const int bk = t & 15, bn = t >> 4;          // B: k bk, columns bn, bn + 16, bn + 32, bn + 48
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = k0 + ak + 4 * i, m = m0 + am;
            ra[i] = (m < M && k < K) ? __ldg(A + (size_t)k * M + m) : 0.f;
            const int kk = k0 + bk, n = n0 + bn + 16 * i;
            rb[i] = (kk < K && n < N) ? __ldg(B + (size_t)n * K + kk) : 0.f;
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += kBK) {
#pragma unroll
        for (int i = 0; i < 4; i++) { As[ak + 4 * i][am] = ra[i]; Bs[bk][bn + 16 * i] = rb[i]; }
        __syncthreads();
        if (k0 + kBK < K) fetch(k0 + kBK);
#pragma unroll
        for (int k = 0; k < kBK; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][4 * tx]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][4 * ty]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int n = n0 + 4 * ty + j;
        if (n >= N) continue;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = m0 + 4 * tx + i;
            if (m < M) C[(size_t)n * M + m] = acc[i][j];
        }
    }
}

struct BatArgs {
    const int *parent, *child_first, *child_count, *omega_idx, *cum;
    const float *prob, *diag, *beta, *uhat, *e, *xcur, *uprev, *uhat_prev;
    const float *sxmin, *sxmax, *sxs, *sumin, *sumax;
    const float *a, *b, *c;                 // hoisted per-node products of k_stream
    float *qbar, *r, *sigma, *xb, *y, *y3;  // q_bar [nodes*nx], r, sigma, X [nodes*nv], Y [nodes*max dim], Y3 [nodes*nu]
    float *V, *U, *X;
    const float *w_xi, *w_psi;
    float *pri_xi, *pri_psi, *dual_xi, *dual_psi;
    double *dist_part;
    int nx, nu, nv, N, cs, K, nodes;
    int df_mode, fuse_prox;
    float inv_step;
};

// ---- backward scans -----------------------------------------------------------------------------------------------------
// chains (stages cs .. N-1): thread = (chain j, element e); q = c + q_child (:651-658), q_bar of the leaf is 0
__global__ void k_bat_q_chain(const BatArgs S) {
    const int nx = S.nx, K = S.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)K * nx) return;
    float run = 0.f;
    for (int s = S.N - 1; s >= S.cs; s--) {
        const size_t o = (size_t)__ldg(S.cum + s) * nx + t;
        S.qbar[o] = run;
        run = __ldg(S.c + o) + run;
    }
}
// one branching stage: q_bar_i = sum over the children of q_c = c_c + q_bar_c, first child first (solveSumChildren,
// Utilities.cu:168-201)
__global__ void k_bat_q_stage(const BatArgs S, int first, int count) {
    const int nx = S.nx;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)count * nx) return;
    const int i = first + (int)(t / nx), e = (int)(t % nx);
    const int c0 = __ldg(S.child_first + i), nc = __ldg(S.child_count + i);
    float s = 0.f;
    for (int c = 0; c < nc; c++) {
        const size_t o = (size_t)(c0 + c) * nx + e;
        const float qc = __ldg(S.c + o) + S.qbar[o];
        s = c == 0 ? qc : s + qc;
    }
    S.qbar[(size_t)i * nx + e] = s;
}
// chains: sigma = beta + r_child (:599), r = (sigma + a) + G q_bar (:631-646), X = -1/2 (sigma + G q_bar)  [df: -1/2 r]
__global__ void k_bat_r_chain(const BatArgs S) {
    const int nv = S.nv, K = S.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)K * nv) return;
    float run = 0.f;
    for (int s = S.N - 1; s >= S.cs; s--) {
        const size_t o = (size_t)__ldg(S.cum + s) * nv + t;
        const float sg = __ldg(S.beta + o) + run, y1 = S.y[o];
        const float rr = (sg + __ldg(S.a + o)) + y1;
        S.sigma[o] = sg; S.r[o] = rr;
        S.xb[o] = -0.5f * (S.df_mode ? rr : sg + y1);
        run = rr;
    }
}
__global__ void k_bat_r_stage(const BatArgs S, int first, int count) {
    const int nv = S.nv;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)count * nv) return;
    const int i = first + (int)(t / nv), e = (int)(t % nv);
    const int c0 = __ldg(S.child_first + i), nc = __ldg(S.child_count + i);
    const size_t o = (size_t)i * nv + e;
    float sg = __ldg(S.beta + o);
    if (nc > 0) {
        float rs = S.r[(size_t)c0 * nv + e];
        for (int c = 1; c < nc; c++) rs += S.r[(size_t)(c0 + c) * nv + e];
        sg += rs;
    }
    const float y1 = S.y[o], rr = (sg + __ldg(S.a + o)) + y1;
    S.sigma[o] = sg; S.r[o] = rr;
    S.xb[o] = -0.5f * (S.df_mode ? rr : sg + y1);
}
// v = (-1/2 Omega sigma + Theta q_bar) + (Phi xi + Psi psi)  (:604-627) = OmegaBar X / p + b  [df: -1/2 Omega r = OmegaBar X / p]
__global__ void k_bat_v(const BatArgs S) {
    const int nv = S.nv;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)S.nodes * nv) return;
    const int i = (int)(t / nv);
    const float ip = 1.f / __ldg(S.prob + __ldg(S.omega_idx + i));     // Omega_i = OmegaBar / p (Engine.cu:210-221, 707-716)
    float v = S.y[t] * ip;
    if (!S.df_mode) v += __ldg(S.b + t);
    S.V[t] = v;
}

// ---- forward scans ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fwd_u(bool branching, float uh, float up, float uhp, float lv) {
    return branching ? (up + -1.f * uhp) + (uh + lv) : ((uh + up) + -1.f * uhp) + lv;      // :701-710 / :683-693, :722-728
}
__global__ void k_bat_u_stage(const BatArgs S, int first, int count, int branching) {
    const int nu = S.nu;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)count * nu) return;
    const int i = first + (int)(t / nu), e = (int)(t % nu);
    const int par = __ldg(S.parent + i);
    const size_t o = (size_t)i * nu + e;
    const float up = par < 0 ? __ldg(S.uprev + e) : S.U[(size_t)par * nu + e];
    const float uhp = par < 0 ? __ldg(S.uhat_prev + e) : __ldg(S.uhat + (size_t)par * nu + e);
    S.U[o] = fwd_u(branching != 0, __ldg(S.uhat + o), up, uhp, S.y3[o]);
}
__global__ void k_bat_u_chain(const BatArgs S, int head_branching) {
    const int nu = S.nu, K = S.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)K * nu) return;
    const int j = (int)(t / nu), e = (int)(t % nu);
    const int par = __ldg(S.parent + __ldg(S.cum + S.cs) + j);
    float up = par < 0 ? __ldg(S.uprev + e) : S.U[(size_t)par * nu + e];
    float uhp = par < 0 ? __ldg(S.uhat_prev + e) : __ldg(S.uhat + (size_t)par * nu + e);
    for (int s = S.cs; s < S.N; s++) {
        const size_t o = (size_t)__ldg(S.cum + s) * nu + t;
        const float uh = __ldg(S.uhat + o);
        const float u = fwd_u(s == S.cs && head_branching, uh, up, uhp, S.y3[o]);
        S.U[o] = u;
        up = u; uhp = uh;
    }
}
__device__ __forceinline__ float fwd_x(bool branching, float xp, float ei, float bu) {
    return branching ? xp + (ei + bu) : (xp + ei) + bu;                                    // :712-719 / :730-737
}
__global__ void k_bat_x_stage(const BatArgs S, int first, int count, int branching) {
    const int nx = S.nx;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)count * nx) return;
    const int i = first + (int)(t / nx), e = (int)(t % nx);
    const int par = __ldg(S.parent + i);
    const size_t o = (size_t)i * nx + e;
    const float xp = par < 0 ? __ldg(S.xcur + e) : S.X[(size_t)par * nx + e];
    S.X[o] = fwd_x(branching != 0, xp, __ldg(S.e + o), S.y[o]);
}
__global__ void k_bat_x_chain(const BatArgs S, int head_branching) {
    const int nx = S.nx, K = S.K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)K * nx) return;
    const int j = (int)(t / nx), e = (int)(t % nx);
    const int par = __ldg(S.parent + __ldg(S.cum + S.cs) + j);
    float xr = par < 0 ? __ldg(S.xcur + e) : S.X[(size_t)par * nx + e];
    for (int s = S.cs; s < S.N; s++) {
        const size_t o = (size_t)__ldg(S.cum + s) * nx + t;
        xr = fwd_x(s == S.cs && head_branching, xr, __ldg(S.e + o), S.y[o]);
        S.X[o] = xr;
    }
}
// Hx = sysF x, Hu = sysG u (:744-747) and, if fused, the box part of proximalFunG (:778-789, :827): one CTA per node, one
// slot of distance partial sums per node
__global__ void __launch_bounds__(256) k_bat_epilogue(const BatArgs S) {
    __shared__ double dsh[16];
    const int nx = S.nx, nu = S.nu, ny = 2 * nx + nu, i = blockIdx.x;
    const float *dg = S.diag + (size_t)i * ny;
    double s1 = 0, s2 = 0;
    for (int t = threadIdx.x; t < ny; t += 256) {
        if (t < 2 * nx) {
            const int j = t < nx ? t : t - nx;
            const size_t k = (size_t)i * 2 * nx + t, kb = (size_t)i * nx + j;
            const float hx = __ldg(dg + t) * S.X[kb];
            S.pri_xi[k] = hx;
            if (S.fuse_prox) {
                const float tt = hx + S.inv_step * S.w_xi[k];
                const float z = t < nx ? clampf(tt, __ldg(S.sxmin + kb), __ldg(S.sxmax + kb)) : clampf(tt, __ldg(S.sxs + kb), __int_as_float(0x7F7F7F7F));
                S.dual_xi[k] = z;
                const float df = tt + -1.f * z;
                if (t < nx) s1 += (double)df * df; else s2 += (double)df * df;
            }
        } else {
            const size_t k = (size_t)i * nu + (t - 2 * nx);
            const float hu = __ldg(dg + t) * S.U[k];
            S.pri_psi[k] = hu;
            if (S.fuse_prox) S.dual_psi[k] = clampf(hu + S.inv_step * S.w_psi[k], __ldg(S.sumin + k), __ldg(S.sumax + k));
        }
    }
    if (!S.fuse_prox) return;
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { dsh[warp] = s1; dsh[8 + warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t1 += dsh[w]; t2 += dsh[8 + w]; }
        S.dist_part[2 * (size_t)i] = t1; S.dist_part[2 * (size_t)i + 1] = t2;
    }
}

static void sgemm(cudaStream_t st, int M, int N, int K, const float *A, const float *B, float *C) {
    k_sgemm<<<dim3(ceil_div(M, kBM), ceil_div(N, kBN)), 256, 0, st>>>(M, N, K, A, B, C);
}

// scratch of this mode, on first use (outside any stream capture)
rn_status batched_prepare(Handle *h) {
    if (h->bat_y) return RN_OK;
    const rn_dims &d = h->d;
    const size_t n = d.nodes, dmax = std::max(std::max(d.nv, d.nu), d.nx);
    RN_CHECK(dev_alloc(h, &h->bat_y, n * dmax)); RN_CHECK(dev_alloc(h, &h->bat_y3, n * d.nu));
    RN_CHECK(dev_alloc(h, &h->bat_x, n * d.nv));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status launch_sweeps_batched(Handle *h, cudaStream_t st, bool fuse_prox, int *n_launch, int *n_slots, cudaEvent_t mid) {
    const rn_dims &d = h->d;
    const size_t n = d.nodes;
    if (!h->bat_y) return fail(h, RN_ERR_STATE, "batched sweeps before batched_prepare");
    int cs = h->chain_stage;   // first stage of the non-branching tail (N: none -- every stage goes stage by stage)
    if (cs < d.N && h->h_nps[cs] != d.K) cs = d.N;
    BatArgs S{};
    S.parent = h->t.parent; S.child_first = h->t.child_first; S.child_count = h->t.child_count; S.omega_idx = h->t.omega_idx;
    S.cum = h->cum_dev; S.prob = h->t.prob; S.diag = h->diag; S.beta = h->beta; S.uhat = h->uhat; S.e = h->e;
    S.xcur = h->xcur; S.uprev = h->uprev; S.uhat_prev = h->uhat_prev;
    S.sxmin = h->sxmin; S.sxmax = h->sxmax; S.sxs = h->sxs; S.sumin = h->sumin; S.sumax = h->sumax;
    S.a = h->a; S.b = h->b; S.c = h->c; S.qbar = h->q; S.r = h->r; S.sigma = h->sigma; S.xb = h->bat_x; S.y = h->bat_y; S.y3 = h->bat_y3;
    S.V = h->V; S.U = h->U; S.X = h->X; S.w_xi = h->acc_xi; S.w_psi = h->acc_psi;
    S.pri_xi = h->pri_xi; S.pri_psi = h->pri_psi; S.dual_xi = h->dual_xi; S.dual_psi = h->dual_psi; S.dist_part = h->dist_part;
    S.nx = d.nx; S.nu = d.nu; S.nv = d.nv; S.N = d.N; S.cs = cs; S.K = d.K; S.nodes = d.nodes;
    S.df_mode = h->factor_mode == RN_FACTORS_DF ? 1 : 0; S.fuse_prox = fuse_prox ? 1 : 0; S.inv_step = 1 / h->step;
    int launches = 0;
    auto blocks = [](size_t work) { return (unsigned)((work + 255) / 256); };
    auto branching = [&](int s) { return s > 0 && h->h_nps[s] > h->h_nps[s - 1] ? 1 : 0; };
    // ---- backward
    if (cs < d.N) { k_bat_q_chain<<<blocks((size_t)d.K * d.nx), 256, 0, st>>>(S); launches++; }
    for (int s = std::min(cs, d.N) - 1; s >= 0; s--) {
        k_bat_q_stage<<<blocks((size_t)h->h_nps[s] * d.nx), 256, 0, st>>>(S, h->h_cum[s], h->h_nps[s]); launches++;
    }
    sgemm(st, d.nv, d.nodes, d.nx, h->G, h->q, h->bat_y); launches++;                    // Y1 = G Q_bar   (:644-646)
    if (cs < d.N) { k_bat_r_chain<<<blocks((size_t)d.K * d.nv), 256, 0, st>>>(S); launches++; }
    for (int s = std::min(cs, d.N) - 1; s >= 0; s--) {
        k_bat_r_stage<<<blocks((size_t)h->h_nps[s] * d.nv), 256, 0, st>>>(S, h->h_cum[s], h->h_nps[s]); launches++;
    }
    sgemm(st, d.nv, d.nodes, d.nv, h->OmegaBar, h->bat_x, h->bat_y); launches++;           // Y2 = OmegaBar X  (:604-613)
    k_bat_v<<<blocks(n * d.nv), 256, 0, st>>>(S); launches++;
    sgemm(st, d.nu, d.nodes, d.nv, h->L, h->V, h->bat_y3); launches++;                     // Y3 = L V   (:701, :727)
    if (mid) RN_CUDA(h, cudaEventRecord(mid, st));
    // ---- forward
    for (int s = 0; s < std::min(cs, d.N); s++) {
        k_bat_u_stage<<<blocks((size_t)h->h_nps[s] * d.nu), 256, 0, st>>>(S, h->h_cum[s], h->h_nps[s], branching(s)); launches++;
    }
    if (cs < d.N) { k_bat_u_chain<<<blocks((size_t)d.K * d.nu), 256, 0, st>>>(S, branching(cs)); launches++; }
    sgemm(st, d.nx, d.nodes, d.nu, h->B, h->U, h->bat_y); launches++;                      // Y4 = B U   (:715, :736)
    for (int s = 0; s < std::min(cs, d.N); s++) {
        k_bat_x_stage<<<blocks((size_t)h->h_nps[s] * d.nx), 256, 0, st>>>(S, h->h_cum[s], h->h_nps[s], branching(s)); launches++;
    }
    if (cs < d.N) { k_bat_x_chain<<<blocks((size_t)d.K * d.nx), 256, 0, st>>>(S, branching(cs)); launches++; }
    k_bat_epilogue<<<d.nodes, 256, 0, st>>>(S); launches++;
    RN_CUDA(h, cudaGetLastError());
    *n_launch = launches;
    *n_slots = d.nodes;
    return RN_OK;
}

}  // namespace rn
