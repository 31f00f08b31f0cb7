// rn_affine.cu -- per-solve affine terms: Engine::updateStateControl and
// Engine::eliminateInputDistubanceCoupling as two node-parallel kernels.
//
// Reference: /root/reference/src/Engine.cu:1300-1316 (uhat_prev = Lhat d_prev) and :1147-1298
// (13 cudaMalloc, K H2D copies of Gd, three host loops of `nodes` Saxpy launches, N+1 batched
// GEMMs); tree kernels /root/reference/src/Utilities.cu:69-131 (calculateDiffUhat, calculateZeta).
//   d_i = d_hat[stage] + err_i ; e_i = Gd d_i ; uhat_i = Lhat d_i
//   alpha_i = w_e (err_price_i + (alpha_hat[stage] + alpha1))
//   dU_0 = uhat_0 - uhat_prev ; dU_i = uhat_i - uhat_parent
//   zeta_i = p_i dU_i - sum_children p_c dU_c
//   beta_i = 2 (W L)' zeta_i + p_i L' alpha_i
#include "rn_internal.h"

namespace rn {

__global__ void k_gemv_small(int m, int n, const float *__restrict__ A, const float *__restrict__ x, float *__restrict__ y) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < n; c++) acc += A[r + (size_t)c * m] * x[c];
        y[r] = acc;
    }
}

__global__ void __launch_bounds__(256)
k_elim_node(int nodes, int nx, int nu, int nd, const int *__restrict__ stages, const float *__restrict__ err_demand,
            const float *__restrict__ err_price, const float *__restrict__ dhat, const float *__restrict__ alphahat,
            const float *__restrict__ alpha1, const float *__restrict__ Gd, const float *__restrict__ Lhat,
            int demand_unc, int price_unc, float w_econ, float *__restrict__ e, float *__restrict__ uhat,
            float *__restrict__ alpha) {
    extern __shared__ float sm[];   // d[nd]
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const int s = stages[i];
    for (int t = threadIdx.x; t < nd; t += blockDim.x)
        sm[t] = (demand_unc ? err_demand[(size_t)i * nd + t] : 0.f) + dhat[(size_t)s * nd + t];
    __syncthreads();
    for (int r = threadIdx.x; r < nx + nu; r += blockDim.x) {
        float acc = 0.f;
        if (r < nx) {
            for (int c = 0; c < nd; c++) acc += Gd[r + (size_t)c * nx] * sm[c];
            e[(size_t)i * nx + r] = acc;
        } else {
            const int rr = r - nx;
            for (int c = 0; c < nd; c++) acc += Lhat[rr + (size_t)c * nu] * sm[c];
            uhat[(size_t)i * nu + rr] = acc;
        }
    }
    for (int t = threadIdx.x; t < nu; t += blockDim.x) {
        const float ah = alphahat[(size_t)s * nu + t] + alpha1[t];
        alpha[(size_t)i * nu + t] = w_econ * ((price_unc ? err_price[(size_t)i * nu + t] : 0.f) + ah);
    }
}

__global__ void __launch_bounds__(256)
k_elim_beta(int nodes, int nu, int nv, const int *__restrict__ parent, const int *__restrict__ child_first,
            const int *__restrict__ child_count, const float *__restrict__ prob, const float *__restrict__ uhat,
            const float *__restrict__ uhat_prev, const float *__restrict__ alpha, const float *__restrict__ Wv,
            const float *__restrict__ L, float *__restrict__ zeta, float *__restrict__ beta) {
    extern __shared__ float sm[];   // zeta[nu] | alpha[nu]
    float *sz = sm, *sa = sm + nu;
    const int i = blockIdx.x;
    if (i >= nodes) return;
    const float p = prob[i];
    const int par = parent[i], c0 = child_first[i], nc = child_count[i];
    for (int t = threadIdx.x; t < nu; t += blockDim.x) {
        const float ui = uhat[(size_t)i * nu + t];
        const float dU = ui - (par < 0 ? uhat_prev[t] : uhat[(size_t)par * nu + t]);
        float z = p * dU;
        for (int c = 0; c < nc; c++) z = z - prob[c0 + c] * (uhat[(size_t)(c0 + c) * nu + t] - ui);
        sz[t] = z;
        zeta[(size_t)i * nu + t] = z;
        sa[t] = alpha[(size_t)i * nu + t];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nv; r += nwarps) {
        float s1 = 0.f, s2 = 0.f;
        for (int t = lane; t < nu; t += 32) {
            s1 += Wv[t + (size_t)r * nu] * sz[t];
            s2 += L[t + (size_t)r * nu] * sa[t];
        }
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) beta[(size_t)i * nv + r] = 2.f * s1 + p * s2;
    }
}

// beta = 2 (W L)' zeta + p L' alpha for nodes [first, first + count) from the zeta rows already in place
__global__ void __launch_bounds__(256)
k_beta_from_zeta(int first, int nu, int nv, const float *__restrict__ prob, const float *__restrict__ zeta,
                 const float *__restrict__ alpha, const float *__restrict__ Wv, const float *__restrict__ L,
                 float *__restrict__ beta) {
    extern __shared__ float sm[];   // zeta[nu] | alpha[nu]
    float *sz = sm, *sa = sm + nu;
    const int i = first + blockIdx.x;
    const float p = prob[i];
    for (int t = threadIdx.x; t < nu; t += blockDim.x) { sz[t] = zeta[(size_t)i * nu + t]; sa[t] = alpha[(size_t)i * nu + t]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nv; r += nwarps) {
        float s1 = 0.f, s2 = 0.f;
        for (int t = lane; t < nu; t += 32) {
            s1 += Wv[t + (size_t)r * nu] * sz[t];
            s2 += L[t + (size_t)r * nu] * sa[t];
        }
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (lane == 0) beta[(size_t)i * nv + r] = 2.f * s1 + p * s2;
    }
}

rn_status fix_beta(Handle *h, int first, int count, const float *zeta_rows) {
    if (count == 0) return RN_OK;
    const rn_dims &d = h->d;
    RN_CUDA(h, cudaMemcpyAsync(h->zeta + (size_t)first * d.nu, zeta_rows, (size_t)count * d.nu * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
    k_beta_from_zeta<<<count, 256, 2 * d.nu * sizeof(float), h->stream>>>(first, d.nu, d.nv, h->t.prob, h->zeta, h->alpha, h->Wv,
                                                                           h->L, h->beta);
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    RN_CUDA(h, cudaStreamSynchronize(h->stream));   // zeta_rows is the caller's buffer
    return RN_OK;
}

rn_status update_state(Handle *h, const float *x, const float *u_prev, const float *d_prev) {
    const rn_dims &d = h->d;
    // staged through pinned memory: one async H2D per vector, no driver-side pageable staging
    RN_CUDA(h, cudaStreamSynchronize(h->stream));   // the pinned block may still feed a previous copy
    float *p = h->pinned;
    memcpy(p, x, d.nx * sizeof(float));
    memcpy(p + d.nx, u_prev, d.nu * sizeof(float));
    memcpy(p + d.nx + d.nu, d_prev, d.nd * sizeof(float));
    RN_CUDA(h, cudaMemcpyAsync(h->xcur, p, d.nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->uprev, p + d.nx, d.nu * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->dprev, p + d.nx + d.nu, d.nd * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    k_gemv_small<<<ceil_div(d.nu, 128), 128, 0, h->stream>>>(d.nu, d.nd, h->Lhat, h->dprev, h->uhat_prev);
    h->launches += 1;
    RN_CUDA(h, cudaGetLastError());
    h->state_set = true;
    return RN_OK;
}

rn_status eliminate_coupling(Handle *h, const float *d_hat, const float *alpha_hat) {
    const rn_dims &d = h->d;
    RN_CUDA(h, cudaStreamSynchronize(h->stream));   // pinned staging may still feed a previous copy
    float *p = h->pinned + (d.nx + d.nu + d.nd);
    const size_t nd_all = (size_t)d.N * d.nd, nu_all = (size_t)d.N * d.nu;
    // (the vectors staged by update_state sit in front of this region and stay valid)
    memcpy(p, d_hat, nd_all * sizeof(float));
    memcpy(p + nd_all, alpha_hat, nu_all * sizeof(float));
    RN_CUDA(h, cudaMemcpyAsync(h->dhat, p, nd_all * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(h->alphahat, p + nd_all, nu_all * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    k_elim_node<<<d.nodes, 256, d.nd * sizeof(float), h->stream>>>(
        d.nodes, d.nx, d.nu, d.nd, h->t.stages, h->t.err_demand, h->t.err_price, h->dhat, h->alphahat, h->alpha1, h->Gd,
        h->Lhat, h->demand_uncertainty ? 1 : 0, h->price_uncertainty ? 1 : 0, h->w_econ, h->e, h->uhat, h->alpha);
    k_elim_beta<<<d.nodes, 256, 2 * d.nu * sizeof(float), h->stream>>>(
        d.nodes, d.nu, d.nv, h->t.parent, h->t.child_first, h->t.child_count, h->t.prob, h->uhat, h->uhat_prev, h->alpha,
        h->Wv, h->L, h->zeta, h->beta);
    h->launches += 2;
    RN_CUDA(h, cudaGetLastError());
    h->eliminated = true;
    return RN_OK;
}

}  // namespace rn
