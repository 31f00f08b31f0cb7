// rn_factor.cu -- Engine::factorStep + initialiseSystemDevice, B200 version.
//
// Reference: /root/reference/src/Engine.cu:382-463 (system init, preconditioning),
// :466-669 (L, Lhat through cuSOLVER Dgesvd), :671-774 (factor step), :1318-1359 (batched
// LU inverse); kernels /root/reference/src/Utilities.cu:33-58, 360-405.
//
// The reference forms every per-node matrix with batched cuBLAS GEMMs against dense
// diagonal matrices.  All of them are (shared matrix) x (stage diagonal) x (1/p):
//     sysF_i = [diag(s_x); diag(s_xs)], sysG_i = diag(s_u),  s = sqrt(p_i) * precond[stage]
//     F_i = L' sysG_i'          = L' . colscale(s_u)
//     D_i = Bbar' sysF_i'       = Bbar' . colscale(s_x | s_xs)
//     Omega_i = (p_i Rbar)^-1   = Rbar^-1 / p_i
//     Theta_i = -1/2 Omega_i Bbar',  Phi_i = -1/2 Omega_i D_i,  Psi_i = -1/2 Omega_i F_i
// so the shared small matrices are built once (host, fp64, then rounded to fp32) and the
// per-node outputs are written by one bandwidth-bound kernel in the reference's packed
// layouts -- those are the buffers the stream kernel reads every APG iteration.
#include <cusolverDn.h>

#include <cmath>

#include "rn_internal.h"

namespace rn {

// ------------------------------------------------------------------------------------------
// host fp64 helpers (matrices are tiny: <= nu x nu)
// ------------------------------------------------------------------------------------------
static bool invert_gauss_jordan(int n, std::vector<double> &A, std::vector<double> &inv) {
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) inv[i + (size_t)i * n] = 1.0;
    for (int k = 0; k < n; k++) {
        int p = k; double mx = std::fabs(A[k + (size_t)k * n]);
        for (int i = k + 1; i < n; i++) { double v = std::fabs(A[i + (size_t)k * n]); if (v > mx) { mx = v; p = i; } }
        if (mx == 0.0) return false;
        if (p != k)
            for (int j = 0; j < n; j++) { std::swap(A[k + (size_t)j * n], A[p + (size_t)j * n]); std::swap(inv[k + (size_t)j * n], inv[p + (size_t)j * n]); }
        double d = 1.0 / A[k + (size_t)k * n];
        for (int j = 0; j < n; j++) { A[k + (size_t)j * n] *= d; inv[k + (size_t)j * n] *= d; }
        for (int i = 0; i < n; i++) {
            if (i == k) continue;
            double f = A[i + (size_t)k * n];
            if (f == 0.0) continue;
            for (int j = 0; j < n; j++) { A[i + (size_t)j * n] -= f * A[k + (size_t)j * n]; inv[i + (size_t)j * n] -= f * inv[k + (size_t)j * n]; }
        }
    }
    return true;
}

// C(m x n) = alpha * op(A) op(B), column-major, fp64 accumulate from fp32/fp64 inputs
template <typename TA, typename TB>
static void hgemm(bool ta, bool tb, int m, int n, int k, double alpha, const TA *A, int lda, const TB *B, int ldb,
                  std::vector<double> &C) {
    C.assign((size_t)m * n, 0.0);
    for (int j = 0; j < n; j++)
        for (int p = 0; p < k; p++) {
            double b = tb ? (double)B[j + (size_t)p * ldb] : (double)B[p + (size_t)j * ldb];
            if (b == 0.0) continue;
            for (int i = 0; i < m; i++) {
                double a = ta ? (double)A[p + (size_t)i * lda] : (double)A[i + (size_t)p * lda];
                C[i + (size_t)j * m] += alpha * a * b;
            }
        }
}

// Engine::calculateMatLandMatLhat (Engine.cu:466-669): the same cuSOLVER routine on the same
// matrix (E' as nu x ne, jobu = jobvt = 'A'), so the null-space basis equals the reference's
// on the same toolkit.  The small products that follow are done on the host in fp64.
static rn_status null_space_svd(Handle *h) {
    const int nu = h->d.nu, ne = h->d.ne, nd = h->d.nd, nv = h->d.nv;
    h->h_L.assign((size_t)nu * nv, 0.f);
    h->h_Lhat.assign((size_t)nu * nd, 0.f);
    if (ne == 0) {   // no mixing nodes: L = I, Lhat = 0
        for (int i = 0; i < nu; i++) h->h_L[i + (size_t)i * nu] = 1.f;
        return RN_OK;
    }
    std::vector<double> Et((size_t)nu * ne);
    for (int r = 0; r < ne; r++) for (int c = 0; c < nu; c++) Et[(size_t)r * nu + c] = (double)h->h_E[r + (size_t)c * ne];
    double *dA = nullptr, *dS = nullptr, *dU = nullptr, *dVT = nullptr, *dWork = nullptr, *dRwork = nullptr;
    int *dInfo = nullptr;
    cusolverDnHandle_t cs = nullptr;
    rn_status rc = RN_OK;
    int lwork = 0, info = 0;
    std::vector<double> U((size_t)nu * nu), S(ne), VT((size_t)ne * ne);
#define SV_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(h, RN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); goto done; } } while (0)
#define SV_CS(call) do { cusolverStatus_t s__ = (call); if (s__ != CUSOLVER_STATUS_SUCCESS) { rc = fail(h, RN_ERR_CUDA, "%s: cusolver status %d", #call, (int)s__); goto done; } } while (0)
    SV_CUDA(cudaMalloc(&dA, sizeof(double) * nu * ne)); SV_CUDA(cudaMalloc(&dS, sizeof(double) * ne));
    SV_CUDA(cudaMalloc(&dU, sizeof(double) * nu * nu)); SV_CUDA(cudaMalloc(&dVT, sizeof(double) * ne * ne));
    SV_CUDA(cudaMalloc(&dInfo, sizeof(int))); SV_CUDA(cudaMalloc(&dRwork, sizeof(double) * (ne < nu ? ne : nu)));
    SV_CUDA(cudaMemcpy(dA, Et.data(), sizeof(double) * nu * ne, cudaMemcpyHostToDevice));
    SV_CS(cusolverDnCreate(&cs));
    SV_CS(cusolverDnSetStream(cs, h->stream));
    SV_CS(cusolverDnDgesvd_bufferSize(cs, nu, ne, &lwork));
    SV_CUDA(cudaMalloc(&dWork, sizeof(double) * (lwork > 0 ? lwork : 1)));
    SV_CS(cusolverDnDgesvd(cs, 'A', 'A', nu, ne, dA, nu, dS, dU, nu, dVT, ne, dWork, lwork, dRwork, dInfo));
    SV_CUDA(cudaMemcpyAsync(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaMemcpyAsync(U.data(), dU, sizeof(double) * nu * nu, cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaMemcpyAsync(S.data(), dS, sizeof(double) * ne, cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaMemcpyAsync(VT.data(), dVT, sizeof(double) * ne * ne, cudaMemcpyDeviceToHost, h->stream));
    SV_CUDA(cudaStreamSynchronize(h->stream));
    if (info != 0) { rc = fail(h, RN_ERR_CUDA, "cusolverDnDgesvd did not converge (info %d)", info); goto done; }
    {
        // pinv(E) = U(:,0:ne) diag(1/S) VT  (Engine.cu:574-582);  Lhat = -pinv(E) Ed (:584-586);  L = U(:, ne:nu) (:613-617)
        std::vector<double> T((size_t)ne * ne), P, Lh;
        for (int c = 0; c < ne; c++) for (int r = 0; r < ne; r++) { double s = S[r]; T[r + (size_t)c * ne] = (std::fabs(s) > 0 ? 1.0 / s : s) * VT[r + (size_t)c * ne]; }
        hgemm(false, false, nu, ne, ne, 1.0, U.data(), nu, T.data(), ne, P);
        hgemm(false, false, nu, nd, ne, -1.0, P.data(), nu, h->h_Ed.data(), ne, Lh);
        for (size_t i = 0; i < Lh.size(); i++) h->h_Lhat[i] = (float)Lh[i];
        for (int c = 0; c < nv; c++) for (int r = 0; r < nu; r++) h->h_L[r + (size_t)c * nu] = (float)U[r + (size_t)(ne + c) * nu];
    }
done:
#undef SV_CUDA
#undef SV_CS
    if (cs) cusolverDnDestroy(cs);
    cudaFree(dA); cudaFree(dS); cudaFree(dU); cudaFree(dVT); cudaFree(dWork); cudaFree(dRwork); cudaFree(dInfo);
    return rc;
}

// ------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------

// preconditionSystem / preconditionConstraintX / preconditionConstraintU (Utilities.cu:33-58, 360-405),
// one CTA per node; also fills devSysXsUpper with 0x7F7F7F7F (Engine.cu:454-455).
__global__ void k_precondition(int nodes, int nx, int nu, const int *__restrict__ stages, const float *__restrict__ prob,
                               const float *__restrict__ precond, const float *__restrict__ xmin,
                               const float *__restrict__ xmax, const float *__restrict__ xsafe,
                               const float *__restrict__ umin, const float *__restrict__ umax, float *__restrict__ diag,
                               float *__restrict__ sxmin, float *__restrict__ sxmax, float *__restrict__ sxs,
                               float *__restrict__ sxs_upper, float *__restrict__ sumin, float *__restrict__ sumax) {
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const int ny = 2 * nx + nu;
    const float sp = sqrtf(prob[i]);
    const float *pc = precond + (size_t)stages[i] * ny;   // per stage [u | x | xsafe]
    float *dg = diag + (size_t)i * ny;                     // per node  [s_x | s_xs | s_u]
    for (int t = threadIdx.x; t < nx; t += blockDim.x) {
        const float scx = sp * pc[nu + t], scs = sp * pc[nu + nx + t];
        dg[t] = scx; dg[nx + t] = scs;
        sxmax[(size_t)i * nx + t] = scx * xmax[t];
        sxmin[(size_t)i * nx + t] = scx * xmin[t];
        sxs[(size_t)i * nx + t] = scs * xsafe[t];
        sxs_upper[(size_t)i * nx + t] = __int_as_float(0x7F7F7F7F);
    }
    for (int t = threadIdx.x; t < nu; t += blockDim.x) {
        const float sc = sp * pc[t];
        dg[2 * nx + t] = sc;
        sumax[(size_t)i * nu + t] = sc * umax[t];
        sumin[(size_t)i * nu + t] = sc * umin[t];
    }
}

// Omega_i = OmegaBar / p_i, Theta_i = ThetaBar / p_i for the distinct nodes i < n_omega
__global__ void k_build_shared_factors(int n_omega, int nv, int nx, const float *__restrict__ prob,
                                       const float *__restrict__ omega_bar, const float *__restrict__ theta_bar,
                                       float *__restrict__ Omega, float *__restrict__ Theta) {
    const int i = blockIdx.x;
    if (i >= n_omega) return;
    const float p = prob[i];
    float *om = Omega + (size_t)i * nv * nv, *th = Theta + (size_t)i * nv * nx;
    for (int t = threadIdx.x; t < nv * nv; t += blockDim.x) om[t] = omega_bar[t] / p;
    for (int t = threadIdx.x; t < nv * nx; t += blockDim.x) th[t] = theta_bar[t] / p;
}

// Per node, packed reference layouts (Engine.cu:201-207): Phi_i, D_i (nv x 2nx), Psi_i, F_i (nv x nu).
__global__ void k_build_node_factors(int nodes, int nv, int nx, int nu, const int *__restrict__ omega_idx,
                                     const float *__restrict__ prob, const float *__restrict__ diag,
                                     const float *__restrict__ G /* Bbar' nv*nx */, const float *__restrict__ Lt /* L' nv*nu */,
                                     const float *__restrict__ theta_bar, const float *__restrict__ psi_bar,
                                     float *__restrict__ Phi, float *__restrict__ Psi, float *__restrict__ D,
                                     float *__restrict__ F) {
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const int ny = 2 * nx + nu;
    const float pa = prob[omega_idx[i]];
    const float *dg = diag + (size_t)i * ny;
    float *phi = Phi + (size_t)i * nv * 2 * nx, *dd = D + (size_t)i * nv * 2 * nx;
    float *psi = Psi + (size_t)i * nv * nu, *ff = F + (size_t)i * nv * nu;
    for (int t = threadIdx.x; t < nv * 2 * nx; t += blockDim.x) {
        const int col = t / nv, r = t - col * nv;
        const int j = col < nx ? col : col - nx;
        const float s = dg[col];                      // s_x for col < nx, s_xs otherwise
        dd[t] = G[r + j * nv] * s;
        phi[t] = (theta_bar[r + j * nv] / pa) * s;
    }
    for (int t = threadIdx.x; t < nv * nu; t += blockDim.x) {
        const int col = t / nv;
        const float s = dg[2 * nx + col];
        ff[t] = Lt[t] * s;
        psi[t] = (psi_bar[t] / pa) * s;
    }
}

// dense sysF (2nx x nx) / sysG (nu x nu) per node, only for the drop-in getters
__global__ void k_dense_sys(int nodes, int nx, int nu, const float *__restrict__ diag, float *__restrict__ sysF,
                            float *__restrict__ sysG) {
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const float *dg = diag + (size_t)i * (2 * nx + nu);
    float *f = sysF + (size_t)i * 2 * nx * nx, *g = sysG + (size_t)i * nu * nu;
    for (int t = threadIdx.x; t < 2 * nx * nx; t += blockDim.x) {
        const int col = t / (2 * nx), row = t - col * 2 * nx;
        f[t] = (row == col) ? dg[col] : (row == nx + col ? dg[nx + col] : 0.f);
    }
    for (int t = threadIdx.x; t < nu * nu; t += blockDim.x) {
        const int col = t / nu, row = t - col * nu;
        g[t] = (row == col) ? dg[2 * nx + col] : 0.f;
    }
}

rn_status factor_step(Handle *h) {
    const rn_dims &d = h->d;
    const int nx = d.nx, nu = d.nu, nv = d.nv, nd = d.nd, nodes = d.nodes;
    (void)nd;
    if (!h->have_null_space) RN_CHECK(null_space_svd(h));
    // shared small matrices, fp64 on the host, rounded once to fp32
    std::vector<double> Wv, Rbar, Gm, Obar, Tbar, Pbar, Lt;
    hgemm(false, false, nu, nv, nu, 1.0, h->h_W.data(), nu, h->h_L.data(), nu, Wv);          // Wv = W L       (Engine.cu:412-414)
    hgemm(true, false, nv, nv, nu, 1.0, h->h_L.data(), nu, Wv.data(), nu, Rbar);             // Rbar = L' Wv   (:415-416)
    hgemm(true, true, nv, nx, nu, 1.0, h->h_L.data(), nu, h->h_B.data(), nx, Gm);            // G = L' B'      (:702-705)
    {
        std::vector<double> tmp = Rbar;
        if (!invert_gauss_jordan(nv, tmp, Obar))
            return fail(h, RN_ERR_SINGULAR, "factor step: Rbar = L'WL is singular (reference: Engine.cu:1334-1353)");
    }
    hgemm(false, false, nv, nx, nv, -0.5, Obar.data(), nv, Gm.data(), nv, Tbar);             // ThetaBar = -1/2 Rbar^-1 Bbar'
    Lt.assign((size_t)nv * nu, 0.0);
    for (int r = 0; r < nv; r++) for (int c = 0; c < nu; c++) Lt[r + (size_t)c * nv] = (double)h->h_L[c + (size_t)r * nu];
    hgemm(false, false, nv, nu, nv, -0.5, Obar.data(), nv, Lt.data(), nv, Pbar);             // PsiBar = -1/2 Rbar^-1 L'
    auto to_f = [](const std::vector<double> &v) { std::vector<float> f(v.size()); for (size_t i = 0; i < v.size(); i++) f[i] = (float)v[i]; return f; };
    std::vector<float> fWv = to_f(Wv), fR = to_f(Rbar), fG = to_f(Gm), fO = to_f(Obar), fT = to_f(Tbar), fP = to_f(Pbar), fLt = to_f(Lt);
    float *d_obar = h->OmegaBar, *d_tbar = h->ThetaBar, *d_lt = h->Lt;
    RN_CHECK(upload(h, h->L, h->h_L.data(), (size_t)nu * nv)); RN_CHECK(upload(h, h->Lhat, h->h_Lhat.data(), (size_t)nu * d.nd));
    RN_CHECK(upload(h, h->Wv, fWv.data(), fWv.size())); RN_CHECK(upload(h, h->Rbar, fR.data(), fR.size()));
    RN_CHECK(upload(h, h->G, fG.data(), fG.size())); RN_CHECK(upload(h, h->PsiBar, fP.data(), fP.size()));
    RN_CHECK(upload(h, d_obar, fO.data(), fO.size())); RN_CHECK(upload(h, d_tbar, fT.data(), fT.size()));
    RN_CHECK(upload(h, d_lt, fLt.data(), fLt.size()));
    k_precondition<<<nodes, 128, 0, h->stream>>>(nodes, nx, nu, h->t.stages, h->t.prob, h->precond, h->xmin, h->xmax, h->xsafe,
                                                 h->umin, h->umax, h->diag, h->sxmin, h->sxmax, h->sxs, h->sxs_upper,
                                                 h->sumin, h->sumax);
    k_build_shared_factors<<<h->n_omega, 256, 0, h->stream>>>(h->n_omega, nv, nx, h->t.prob, d_obar, d_tbar, h->Omega, h->Theta);
    k_build_node_factors<<<nodes, 512, 0, h->stream>>>(nodes, nv, nx, nu, h->t.omega_idx, h->t.prob, h->diag, h->G, d_lt,
                                                       d_tbar, h->PsiBar, h->Phi, h->Psi, h->D, h->F);
    h->launches += 3;
    RN_CUDA(h, cudaGetLastError());
    RN_CUDA(h, cudaStreamSynchronize(h->stream));   // host vectors above go out of scope
    h->factored = true;
    h->pack_dirty = true;   // the persistent kernel's copy of G, OmegaBar, L, B, L' (rn_persist.cu: refresh_pack)
    h->state_set = false; h->eliminated = false; h->have_duals = false;
    return RN_OK;
}

rn_status materialise_dense_sys(Handle *h) {
    const rn_dims &d = h->d;
    if (!h->sysF_dense) {
        RN_CHECK(dev_alloc(h, &h->sysF_dense, (size_t)d.nodes * 2 * d.nx * d.nx, false));
        RN_CHECK(dev_alloc(h, &h->sysG_dense, (size_t)d.nodes * d.nu * d.nu, false));
    }
    k_dense_sys<<<d.nodes, 256, 0, h->stream>>>(d.nodes, d.nx, d.nu, h->diag, h->sysF_dense, h->sysG_dense);
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    return RN_OK;
}

}  // namespace rn
