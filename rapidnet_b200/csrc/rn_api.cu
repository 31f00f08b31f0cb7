// rn_api.cu -- C-ABI entry points: lifecycle, buffers, thin wrappers (include/rapidnet_b200.h).
#include <algorithm>
#include <cmath>

#include "rn_internal.h"

namespace rn {
thread_local std::string g_create_error;

static bool build_tree_host(Handle *h, const rn_tree *t) {
    const rn_dims &d = h->d;
    h->h_stages.assign(t->stages, t->stages + d.nodes);
    h->h_nps.assign(t->nodes_per_stage, t->nodes_per_stage + d.N + 1);
    h->h_cum.assign(t->nodes_per_stage_cumul, t->nodes_per_stage_cumul + d.N + 2);
    h->h_prob.assign(t->prob, t->prob + d.nodes);
    h->h_parent.resize(d.nodes);
    h->h_child_first.assign(d.nodes, 0);
    h->h_child_count.assign(d.nodes, 0);
    for (int i = 0; i < d.nodes; i++) h->h_parent[i] = t->ancestor[i] - 1;
    for (int i = 0; i < d.n_nonleaf; i++) {   // Utilities.cu:113-128
        int first = (i == 0) ? 1 : t->n_children_cumul[i - 1] + 1;
        int cnt = (i == 0) ? t->n_children_cumul[0] : t->n_children_cumul[i] - t->n_children_cumul[i - 1];
        h->h_child_first[i] = first;
        h->h_child_count[i] = cnt;
        for (int c = 0; c < cnt; c++)
            if (first + c >= d.nodes || h->h_parent[first + c] != i) return false;
    }
    if (h->h_cum[0] != 0 || h->h_cum[d.N] != d.nodes) return false;
    for (int s = 0; s < d.N; s++) {
        if (h->h_cum[s + 1] - h->h_cum[s] != h->h_nps[s]) return false;
        for (int j = 0; j < h->h_nps[s]; j++) if (h->h_stages[h->h_cum[s] + j] != s) return false;
    }
    // ScenarioTree::getFinalBranchNode / getFinalBranchStage (ScenarioTree.cu:149-169)
    h->fb_node = 0; h->fb_stage = 0;
    for (int s = 0; s < d.N - 1; s++)
        if (h->h_nps[s] == h->h_nps[s + 1]) { h->fb_node = h->h_cum[s + 1]; h->fb_stage = s; break; }
    const int hint = d.chain_stage_hint;
    if (hint > 0) {   // one rank's part of a larger tree: the larger tree's first repeated stage is its chain stage
        if (hint >= d.N) return false;
        if (hint < d.N - 1) { h->fb_node = h->h_cum[hint + 1]; h->fb_stage = hint; } else { h->fb_node = 0; h->fb_stage = 0; }
    }
    h->n_omega = h->fb_node > 0 ? h->fb_node : d.nodes;
    h->h_omega_idx.resize(d.nodes);
    for (int s = 0; s < d.N; s++)
        for (int j = 0; j < h->h_nps[s]; j++) {   // Engine.cu:210-221
            int i = h->h_cum[s] + j;
            h->h_omega_idx[i] = (h->fb_node > 0 && h->fb_node <= h->h_cum[s]) ? h->fb_node - d.K + j : i;
            if (h->h_omega_idx[i] < 0 || h->h_omega_idx[i] >= h->n_omega) return false;
        }
    // non-branching tail: every node of stage s' > chain_stage is the only child of node (s'-1, same j)
    int cs = d.N - 1;
    while (cs > 0) {
        int s = cs;   // can stage s-1 join the tail?  needs nps equal and 1:1 parents for stage s
        if (h->h_nps[s] != h->h_nps[s - 1]) break;
        bool ok = true;
        for (int j = 0; j < h->h_nps[s] && ok; j++) ok = (h->h_parent[h->h_cum[s] + j] == h->h_cum[s - 1] + j);
        if (!ok) break;
        cs--;
    }
    if (hint > 0) {
        if (cs > hint) return false;   // the tail must be non-branching from the hinted stage on
        cs = hint;
    }
    // a chain CTA keeps one Omega/Theta in shared memory: the alias must be constant along each chain
    for (int s = cs + 1; s < d.N && cs < d.N; s++)
        for (int j = 0; j < h->h_nps[s]; j++)
            if (h->h_omega_idx[h->h_cum[s] + j] != h->h_omega_idx[h->h_cum[cs] + j]) { cs = d.N; break; }
    h->chain_stage = cs;
    return true;
}

static rn_status allocate(Handle *h) {
    const rn_dims &d = h->d;
    const size_t n = d.nodes, nx = d.nx, nu = d.nu, nv = d.nv, nd = d.nd, ny = 2 * nx + nu;
    DevTree &t = h->t;
    RN_CHECK(dev_alloc(h, &t.stages, n)); RN_CHECK(dev_alloc(h, &t.parent, n));
    RN_CHECK(dev_alloc(h, &t.child_first, n)); RN_CHECK(dev_alloc(h, &t.child_count, n));
    RN_CHECK(dev_alloc(h, &t.omega_idx, n)); RN_CHECK(dev_alloc(h, &t.prob, n));
    RN_CHECK(dev_alloc(h, &t.err_demand, n * nd)); RN_CHECK(dev_alloc(h, &t.err_price, n * nu));
    RN_CHECK(dev_alloc(h, &h->B, nx * nu)); RN_CHECK(dev_alloc(h, &h->Gd, nx * nd));
    RN_CHECK(dev_alloc(h, &h->L, nu * nv)); RN_CHECK(dev_alloc(h, &h->Lhat, nu * nd));
    RN_CHECK(dev_alloc(h, &h->W, nu * nu)); RN_CHECK(dev_alloc(h, &h->Wv, nu * nv));
    RN_CHECK(dev_alloc(h, &h->Rbar, nv * nv)); RN_CHECK(dev_alloc(h, &h->G, nv * nx));
    RN_CHECK(dev_alloc(h, &h->PsiBar, nv * nu)); RN_CHECK(dev_alloc(h, &h->Lt, nv * nu));
    RN_CHECK(dev_alloc(h, &h->OmegaBar, nv * nv)); RN_CHECK(dev_alloc(h, &h->ThetaBar, nv * nx));
    RN_CHECK(dev_alloc(h, &h->precond, (size_t)d.N * ny)); RN_CHECK(dev_alloc(h, &h->alpha1, nu));
    RN_CHECK(dev_alloc(h, &h->xmin, nx)); RN_CHECK(dev_alloc(h, &h->xmax, nx)); RN_CHECK(dev_alloc(h, &h->xsafe, nx));
    RN_CHECK(dev_alloc(h, &h->umin, nu)); RN_CHECK(dev_alloc(h, &h->umax, nu));
    {   // the four per-node factor arrays in one slab, D | F | Phi | Psi (256-byte aligned pieces); D, F first --
        // RN_FACTORS_DF streams only those
        const size_t cxi = (n * nv * 2 * nx + 63) & ~size_t(63), cpsi = (n * nv * nu + 63) & ~size_t(63);
        RN_CHECK(dev_alloc(h, &h->factor_slab, 2 * (cxi + cpsi), false));
        h->D = h->factor_slab; h->F = h->D + cxi; h->Phi = h->F + cpsi; h->Psi = h->Phi + cxi;
        h->factor_df_bytes = (cxi + cpsi) * sizeof(float); h->factor_full_bytes = 2 * (cxi + cpsi) * sizeof(float);
    }
    RN_CHECK(dev_alloc(h, &h->Omega, (size_t)h->n_omega * nv * nv)); RN_CHECK(dev_alloc(h, &h->Theta, (size_t)h->n_omega * nv * nx));
    RN_CHECK(dev_alloc(h, &h->diag, n * ny));
    RN_CHECK(dev_alloc(h, &h->sxmin, n * nx)); RN_CHECK(dev_alloc(h, &h->sxmax, n * nx)); RN_CHECK(dev_alloc(h, &h->sxs, n * nx));
    RN_CHECK(dev_alloc(h, &h->sxs_upper, n * nx)); RN_CHECK(dev_alloc(h, &h->sumin, n * nu)); RN_CHECK(dev_alloc(h, &h->sumax, n * nu));
    RN_CHECK(dev_alloc(h, &h->xcur, nx)); RN_CHECK(dev_alloc(h, &h->uprev, nu)); RN_CHECK(dev_alloc(h, &h->dprev, nd));
    RN_CHECK(dev_alloc(h, &h->uhat_prev, nu));
    RN_CHECK(dev_alloc(h, &h->dhat, (size_t)d.N * nd)); RN_CHECK(dev_alloc(h, &h->alphahat, (size_t)d.N * nu));
    RN_CHECK(dev_alloc(h, &h->e, n * nx)); RN_CHECK(dev_alloc(h, &h->uhat, n * nu)); RN_CHECK(dev_alloc(h, &h->alpha, n * nu));
    RN_CHECK(dev_alloc(h, &h->beta, n * nv)); RN_CHECK(dev_alloc(h, &h->zeta, n * nu));
    RN_CHECK(dev_alloc(h, &h->X, n * nx)); RN_CHECK(dev_alloc(h, &h->U, n * nu)); RN_CHECK(dev_alloc(h, &h->V, n * nv));
    RN_CHECK(dev_alloc(h, &h->sigma, n * nv));
    {   // slab: yA | yB | wA | Hx | z | wB  (xi part, psi part each), 256-B aligned pieces
        const size_t cxi = (n * 2 * nx + 63) & ~size_t(63), cpsi = (n * nu + 63) & ~size_t(63);
        const size_t total = 6 * (cxi + cpsi);
        RN_CHECK(dev_alloc(h, &h->apg_slab, total));
        h->apg_slab_bytes = total * sizeof(float);
        float *p = h->apg_slab;
        h->yA_xi = p; p += cxi; h->yA_psi = p; p += cpsi; h->yB_xi = p; p += cxi; h->yB_psi = p; p += cpsi;
        h->acc_xi = p; p += cxi; h->acc_psi = p; p += cpsi; h->pri_xi = p; p += cxi; h->pri_psi = p; p += cpsi;
        h->dual_xi = p; p += cxi; h->dual_psi = p; p += cpsi;
        h->wA_xi = h->acc_xi; h->wA_psi = h->acc_psi; h->wB_xi = p; p += cxi; h->wB_psi = p; p += cpsi;
        h->upd_xi = h->yA_xi; h->upd_psi = h->yA_psi; h->xi = h->yB_xi; h->psi = h->yB_psi;
    }
    RN_CHECK(dev_alloc(h, &h->res_xi, n * 2 * nx)); RN_CHECK(dev_alloc(h, &h->res_psi, n * nu));
    RN_CHECK(dev_alloc(h, &h->cum_dev, (size_t)d.N + 2));
    RN_CHECK(dev_alloc(h, &h->control_action, nu)); RN_CHECK(dev_alloc(h, &h->state_update, nx));
    RN_CHECK(dev_alloc(h, &h->a, n * nv)); RN_CHECK(dev_alloc(h, &h->b, n * nv)); RN_CHECK(dev_alloc(h, &h->c, n * nx));
    RN_CHECK(dev_alloc(h, &h->q, n * nx)); RN_CHECK(dev_alloc(h, &h->r, n * nv));
    h->dist_slots = std::max(d.nodes, 1024);
    RN_CHECK(dev_alloc(h, &h->dist_part, 2 * (size_t)h->dist_slots));
    RN_CHECK(dev_alloc(h, &h->scal, 16));
    RN_CHECK(dev_alloc(h, &h->iter_dev, 4)); RN_CHECK(dev_alloc(h, &h->done_ctr, 8));
    h->lambda_cap = std::max(h->max_iter, 1) + 8;
    RN_CHECK(dev_alloc(h, &h->lambda_tab, h->lambda_cap)); RN_CHECK(dev_alloc(h, &h->pinf, h->lambda_cap));
    h->pinf_slots = 4096;
    RN_CHECK(dev_alloc(h, &h->pinf_part, 6 * (size_t)h->pinf_slots));
    h->pinned_floats = std::max<size_t>({(size_t)d.N * (nd + nu) + nx + nu + nd + 64, (size_t)h->lambda_cap + 64});
    RN_CUDA(h, cudaMallocHost((void **)&h->pinned, h->pinned_floats * sizeof(float)));
    return RN_OK;
}
}  // namespace rn

using rn::Handle;

extern "C" {

rn_status rn_create(const rn_dims *dims, const rn_tree *tree, const rn_network *net, const rn_config *cfg,
                    int device, rn_handle **out) {
    if (!dims || !tree || !net || !cfg || !out) return rn::fail(nullptr, RN_ERR_INVALID, "rn_create: null argument");
    *out = nullptr;
    const rn_dims &d = *dims;
    if (d.nx <= 0 || d.nu <= 0 || d.nd <= 0 || d.nv <= 0 || d.N <= 0 || d.nodes <= 0 || d.K <= 0 || d.ne < 0)
        return rn::fail(nullptr, RN_ERR_INVALID, "rn_create: non-positive dimension");
    if (d.nv != d.nu - d.ne) return rn::fail(nullptr, RN_ERR_INVALID, "rn_create: nv (%d) != nu - ne (%d)", d.nv, d.nu - d.ne);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return rn::fail(nullptr, RN_ERR_CUDA, "rn_create: no CUDA device (%s); this library has no CPU path",
                        ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return rn::fail(nullptr, RN_ERR_INVALID, "rn_create: device %d of %d", device, ndev);
    Handle *h = new Handle();
    h->d = d;
    h->device = device;
    auto bail = [&](rn_status s) { rn::g_create_error = h->err; rn_destroy(reinterpret_cast<rn_handle *>(h)); return s; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(rn::fail(h, RN_ERR_CUDA, "cudaSetDevice(%d) failed", device));
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(rn::fail(h, RN_ERR_CUDA, "cudaGetDeviceProperties failed"));
    h->sm_count = prop.multiProcessorCount;
    h->l2_bytes = prop.l2CacheSize;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(rn::fail(h, RN_ERR_CUDA, "cudaStreamCreate failed"));
    h->own_stream = true;
    if (cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(rn::fail(h, RN_ERR_CUDA, "cudaStreamCreate failed"));
    if (!rn::build_tree_host(h, tree)) return bail(rn::fail(h, RN_ERR_INVALID, "rn_create: inconsistent scenario tree arrays"));
    const size_t nx = d.nx, nu = d.nu, nd = d.nd, ne = d.ne;
    h->h_B.assign(net->B, net->B + nx * nu); h->h_Gd.assign(net->Gd, net->Gd + nx * nd);
    h->h_E.assign(net->E, net->E + ne * nu); h->h_Ed.assign(net->Ed, net->Ed + ne * nd);
    h->h_xmin.assign(net->xmin, net->xmin + nx); h->h_xmax.assign(net->xmax, net->xmax + nx);
    h->h_xsafe.assign(net->xsafe, net->xsafe + nx); h->h_umin.assign(net->umin, net->umin + nu);
    h->h_umax.assign(net->umax, net->umax + nu); h->h_alpha1.assign(net->alpha1, net->alpha1 + nu);
    h->h_W.assign(cfg->costW, cfg->costW + nu * nu);
    h->h_precond.assign(cfg->precond, cfg->precond + (size_t)d.N * (nu + 2 * nx));
    h->pen_x = cfg->penalty_x; h->pen_xs = cfg->penalty_xs; h->step = cfg->step_size;
    h->w_econ = cfg->weight_economical; h->max_iter = cfg->max_iterations;
    if (!(h->step > 0)) return bail(rn::fail(h, RN_ERR_INVALID, "rn_create: step size must be > 0"));
    rn_status s = rn::allocate(h);
    if (s != RN_OK) return bail(s);
    rn::DevTree &t = h->t;
    bool ok = true;
    ok &= rn::upload(h, t.stages, h->h_stages.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, t.parent, h->h_parent.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, t.child_first, h->h_child_first.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, t.child_count, h->h_child_count.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, t.omega_idx, h->h_omega_idx.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, h->cum_dev, h->h_cum.data(), (size_t)d.N + 2) == RN_OK;
    ok &= rn::upload(h, t.prob, h->h_prob.data(), d.nodes) == RN_OK;
    ok &= rn::upload(h, t.err_demand, tree->err_demand, (size_t)d.nodes * nd) == RN_OK;
    ok &= rn::upload(h, t.err_price, tree->err_price, (size_t)d.nodes * nu) == RN_OK;
    ok &= rn::upload(h, h->B, h->h_B.data(), nx * nu) == RN_OK;
    ok &= rn::upload(h, h->Gd, h->h_Gd.data(), nx * nd) == RN_OK;
    ok &= rn::upload(h, h->W, h->h_W.data(), nu * nu) == RN_OK;
    ok &= rn::upload(h, h->precond, h->h_precond.data(), h->h_precond.size()) == RN_OK;
    ok &= rn::upload(h, h->alpha1, h->h_alpha1.data(), nu) == RN_OK;
    ok &= rn::upload(h, h->xmin, h->h_xmin.data(), nx) == RN_OK;
    ok &= rn::upload(h, h->xmax, h->h_xmax.data(), nx) == RN_OK;
    ok &= rn::upload(h, h->xsafe, h->h_xsafe.data(), nx) == RN_OK;
    ok &= rn::upload(h, h->umin, h->h_umin.data(), nu) == RN_OK;
    ok &= rn::upload(h, h->umax, h->h_umax.data(), nu) == RN_OK;
    if (!ok) return bail(RN_ERR_CUDA);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(rn::fail(h, RN_ERR_CUDA, "upload failed"));
    *out = reinterpret_cast<rn_handle *>(h);
    return RN_OK;
}

rn_status rn_destroy(rn_handle *hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    rn::apg_release_graph(h);
    for (int r = 0; r < 8; r++) if (h->xchg_opened[r] && h->xchg_peer[r]) cudaIpcCloseMemHandle(h->xchg_peer[r]);
    for (void *p : h->allocs) cudaFree(p);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    delete h;
    return RN_OK;
}

const char *rn_last_error(const rn_handle *hh) {
    const Handle *h = reinterpret_cast<const Handle *>(hh);
    return h ? h->err.c_str() : rn::g_create_error.c_str();
}

rn_status rn_set_stream(rn_handle *hh, void *cuda_stream) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    rn::apg_release_graph(h);
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return RN_OK;
}

rn_status rn_get_stream(rn_handle *hh, void **cuda_stream) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !cuda_stream) return RN_ERR_INVALID;
    *cuda_stream = reinterpret_cast<void *>(h->stream);
    return RN_OK;
}

rn_status rn_sync(rn_handle *hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_set_modes(rn_handle *hh, rn_sweep_mode sweep, rn_factor_mode factors) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    if ((sweep != RN_SWEEP_PER_STAGE && sweep != RN_SWEEP_CHAIN && sweep != RN_SWEEP_PERSISTENT && sweep != RN_SWEEP_BATCHED) || (factors != RN_FACTORS_FULL && factors != RN_FACTORS_DF && factors != RN_FACTORS_SHARED))
        return rn::fail(h, RN_ERR_INVALID, "rn_set_modes: unknown mode");
    h->sweep_mode = sweep;
    h->factor_mode = factors;
    return RN_OK;
}

rn_status rn_set_grid_limit(rn_handle *hh, int max_ctas) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    if (max_ctas < 0) return rn::fail(h, RN_ERR_INVALID, "rn_set_grid_limit: max_ctas = %d", max_ctas);
    if (h->dist_world > 1 && max_ctas > 0) return rn::fail(h, RN_ERR_STATE, "rn_set_grid_limit: the tree partition uses the whole GPU");
    h->grid_limit = max_ctas;
    return RN_OK;
}

rn_status rn_set_warm_start(rn_handle *hh, int on) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    h->warm_start = on != 0;
    if (!on) h->have_duals = false;
    return RN_OK;
}

rn_status rn_get_info(rn_handle *hh, rn_info *info) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !info) return RN_ERR_INVALID;
    const rn_dims &d = h->d;
    memset(info, 0, sizeof(*info));
    info->device = h->device; info->sm_count = h->sm_count;
    info->final_branch_node = h->fb_node; info->final_branch_stage = h->fb_stage;
    info->chain_first_stage = h->chain_stage; info->num_omega = h->n_omega;
    info->sweep_mode = h->sweep_mode; info->factor_mode = h->factor_mode;
    info->kernel_launches = h->launches; info->launches_per_iteration = h->launches_per_iter;
    info->device_bytes = h->device_bytes;
    info->factor_bytes = (size_t)d.nodes * d.nv * (4 * (size_t)d.nx + 2 * (size_t)d.nu) * sizeof(float);
    info->stream_bytes_per_iteration = rn::stream_bytes_per_iteration(h);
    info->apg_bytes_per_iteration = rn::apg_bytes_per_iteration(h);
    info->last_stream_ms = h->last_stream_ms;
    if (h->scal) {
        float sc[2] = {0, 0};
        if (cudaMemcpyAsync(sc, h->scal, sizeof(sc), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess &&
            cudaStreamSynchronize(h->stream) == cudaSuccess) {
            info->last_distance_x = sc[0]; info->last_distance_xs = sc[1];
        }
    }
    return RN_OK;
}

rn_status rn_set_null_space(rn_handle *hh, const float *L, const float *Lhat) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !L || !Lhat) return RN_ERR_INVALID;
    h->h_L.assign(L, L + (size_t)h->d.nu * h->d.nv);
    h->h_Lhat.assign(Lhat, Lhat + (size_t)h->d.nu * h->d.nd);
    h->have_null_space = true;
    return RN_OK;
}

rn_status rn_factor_step(rn_handle *hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::factor_step(h);
}

rn_status rn_update_state(rn_handle *hh, const float *x, const float *u_prev, const float *d_prev) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !x || !u_prev || !d_prev) return RN_ERR_INVALID;
    if (!h->factored) return rn::fail(h, RN_ERR_STATE, "rn_update_state before rn_factor_step");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::update_state(h, x, u_prev, d_prev);
}

rn_status rn_eliminate_coupling(rn_handle *hh, const float *d_hat, const float *alpha_hat) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !d_hat || !alpha_hat) return RN_ERR_INVALID;
    if (!h->factored || !h->state_set) return rn::fail(h, RN_ERR_STATE, "rn_eliminate_coupling before factor step / update state");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::eliminate_coupling(h, d_hat, alpha_hat);
}

rn_status rn_set_uncertainty(rn_handle *hh, int demand, int price) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    h->demand_uncertainty = demand != 0;
    h->price_uncertainty = price != 0;
    return RN_OK;
}

rn_status rn_apg_init(rn_handle *hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::apg_init(h);
}

rn_status rn_step(rn_handle *hh, rn_step_kind kind, float lambda) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    // extrapolation, residual and dual update only touch the dual vectors (the reference's tests run the extrapolation on a
    // controller that has not been factored yet, TestSmpcController.cu:114-165); the solve step and the prox need the factor step
    if (!h->factored && (kind == RN_STEP_SOLVE || kind == RN_STEP_PROX)) return rn::fail(h, RN_ERR_STATE, "rn_step before rn_factor_step");
    if (kind == RN_STEP_SOLVE && !h->eliminated) return rn::fail(h, RN_ERR_STATE, "solve step before rn_eliminate_coupling");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::apg_step(h, kind, lambda);
}

rn_status rn_apg_solve(rn_handle *hh, int iterations, float *u0_host, float *primal_infs_host) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || iterations < 0) return RN_ERR_INVALID;
    if (!h->factored || !h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_apg_solve before factor step / eliminate");
    RN_CUDA(h, cudaSetDevice(h->device));
    RN_CHECK(rn::apg_enqueue(h, iterations));
    if (u0_host) RN_CUDA(h, cudaMemcpyAsync(u0_host, h->U, h->d.nu * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (primal_infs_host && iterations > 0)
        RN_CUDA(h, cudaMemcpyAsync(primal_infs_host, h->pinf, iterations * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (u0_host || primal_infs_host) RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_apg_continue(rn_handle *hh, int iterations, const float *lambda_host, int warm_restart) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || iterations < 0) return RN_ERR_INVALID;
    if (!h->factored || !h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_apg_continue before factor step / eliminate");
    RN_CUDA(h, cudaSetDevice(h->device));
    if (warm_restart) return rn::apg_warm(h, iterations);
    return rn::apg_continue(h, iterations, lambda_host);
}

rn_status rn_control_action(rn_handle *hh, const float *x, const float *u_prev, const float *d_prev,
                            const float *d_hat, const float *alpha_hat, int iterations, int clamp, float *u0_host) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !x || !u_prev || !d_prev || !d_hat || !alpha_hat || !u0_host || iterations < 0) return RN_ERR_INVALID;
    if (!h->factored) return rn::fail(h, RN_ERR_STATE, "rn_control_action before rn_factor_step");
    RN_CUDA(h, cudaSetDevice(h->device));
    RN_CHECK(rn::update_state(h, x, u_prev, d_prev));
    RN_CHECK(rn::eliminate_coupling(h, d_hat, alpha_hat));
    RN_CHECK(rn::apg_enqueue(h, iterations));
    const float *src = h->U;
    // devControlAction <- devVecU[0:nu] (:1647); clamped only by the fstream variant (:1649)
    RN_CUDA(h, cudaMemcpyAsync(h->control_action, h->U, h->d.nu * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    if (clamp) { RN_CHECK(rn::clamp_control(h)); src = h->control_action; }
    RN_CUDA(h, cudaMemcpyAsync(u0_host, src, h->d.nu * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_move_forward(rn_handle *hh, float *x_next_host, float *u_applied_host) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !x_next_host || !u_applied_host) return RN_ERR_INVALID;
    if (!h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_move_forward before a solve");
    RN_CUDA(h, cudaSetDevice(h->device));
    RN_CHECK(rn::move_forward(h));
    RN_CUDA(h, cudaMemcpyAsync(x_next_host, h->state_update, h->d.nx * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RN_CUDA(h, cudaMemcpyAsync(u_applied_host, h->control_action, h->d.nu * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_dist_prepare(rn_handle *hh, int world, int rank, int K_global, int chain_offset, const int *head_lo,
                          const int *head_hi, unsigned char *ipc_handle_out) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || !ipc_handle_out)
        return rn::fail(h, RN_ERR_INVALID, "rn_dist_prepare: world %d / rank %d out of range (1..8 ranks)", world, rank);
    if (h->persist_ready || h->xchg) return rn::fail(h, RN_ERR_STATE, "rn_dist_prepare must precede the first solve");
    if (!rn::persistent_supported(h)) return rn::fail(h, RN_ERR_INVALID, "the tree partition needs the persistent kernel, which does not fit this problem");
    const int n_crown = h->h_cum[h->chain_stage];
    if (chain_offset < 0 || chain_offset + h->d.K > K_global) return rn::fail(h, RN_ERR_INVALID, "rn_dist_prepare: chain range outside K_global");
    if (world > 1 && (!head_lo || !head_hi)) return rn::fail(h, RN_ERR_INVALID, "rn_dist_prepare: head ranges missing");
    if (world > 1) {
        // the partition must be aligned to the bottom-crown nodes (stage cs-1): the chains below one of them live on ONE rank
        // (its sum over them, S_p, and its zeta are then formed by that rank alone, in the order of the single-GPU solve)
        const int cs = h->chain_stage;
        if (cs <= 0) return rn::fail(h, RN_ERR_INVALID, "rn_dist_prepare: the tree has no crown to cut below");
        for (int p = h->h_cum[cs - 1]; p < h->h_cum[cs]; p++) {
            const int local = h->h_child_count[p], global = head_hi[p] - head_lo[p];
            if (local != 0 && local != global)
                return rn::fail(h, RN_ERR_INVALID, "rn_dist_prepare: the chains below crown node %d are split across ranks (%d of %d here)", p, local, global);
        }
    }
    h->dist_world = world; h->dist_rank = rank; h->dist_K_glob = K_global; h->dist_chain_off = chain_offset;
    if (world > 1) { h->dist_head_lo.assign(head_lo, head_lo + n_crown); h->dist_head_hi.assign(head_hi, head_hi + n_crown); }
    RN_CUDA(h, cudaSetDevice(h->device));
    RN_CHECK(rn::ensure_xchg(h));
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == RN_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    RN_CUDA(h, cudaIpcGetMemHandle(&mh, h->xchg));
    memcpy(ipc_handle_out, &mh, sizeof(mh));
    return RN_OK;
}

rn_status rn_dist_connect(rn_handle *hh, const unsigned char *handles) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !handles) return RN_ERR_INVALID;
    if (!h->xchg) return rn::fail(h, RN_ERR_STATE, "rn_dist_connect before rn_dist_prepare");
    RN_CUDA(h, cudaSetDevice(h->device));
    for (int r = 0; r < h->dist_world; r++) {
        if (r == h->dist_rank || h->xchg_peer[r]) continue;
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handles + (size_t)r * RN_IPC_HANDLE_BYTES, sizeof(mh));
        void *p = nullptr;
        RN_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->xchg_peer[r] = p; h->xchg_opened[r] = true;
    }
    return RN_OK;
}

rn_status rn_dist_fix_crown_beta(rn_handle *hh, int first, int count, const float *zeta_rows) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !zeta_rows || first < 0 || count < 0 || first + count > h->d.nodes) return RN_ERR_INVALID;
    if (!h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_dist_fix_crown_beta before rn_eliminate_coupling");
    return rn::fix_beta(h, first, count, zeta_rows);
}

rn_status rn_prepare(rn_handle *hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    RN_CUDA(h, cudaSetDevice(h->device));
    if (h->sweep_mode == RN_SWEEP_PERSISTENT && rn::persistent_supported(h)) RN_CHECK(rn::persistent_prepare(h));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_dist_sync_crown_beta(rn_handle *hh, int pull) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return RN_ERR_INVALID;
    if (!h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_dist_sync_crown_beta before rn_eliminate_coupling");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::dist_crown_beta(h, pull);
}

rn_status rn_read_pinf_parts(rn_handle *hh, int iterations, float *host) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !host || iterations < 0) return RN_ERR_INVALID;
    if (iterations > h->pinf4_cap) return rn::fail(h, RN_ERR_STATE, "rn_read_pinf_parts: only %d iterations were logged", h->pinf4_cap);
    // rows 0 .. n-2 come from the persistent kernel, row n-1 from k_finalize.  With a partitioned tree every row holds THIS
    // rank's arg-max (rn_apg_solve's primal_infs_host is then rank-local too); partition.merge_pinf combines the ranks.
    RN_CUDA(h, cudaMemcpyAsync(host, h->pinf4, (size_t)iterations * 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_dist_error(rn_handle *hh, int *timed_out) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !timed_out) return RN_ERR_INVALID;
    *timed_out = 0;
    if (!h->xchg) return RN_OK;
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    const size_t off = rn::xchg_err_offset(h);
    RN_CUDA(h, cudaMemcpy(timed_out, static_cast<char *>(h->xchg) + off, sizeof(int), cudaMemcpyDeviceToHost));
    return RN_OK;
}

rn_status rn_buffer(rn_handle *hh, rn_buffer_id id, void **dev_ptr, size_t *bytes) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !dev_ptr || !bytes) return RN_ERR_INVALID;
    const rn_dims &d = h->d;
    const size_t n = d.nodes, nx = d.nx, nu = d.nu, nv = d.nv, nd = d.nd;
    float *p = nullptr; size_t cnt = 0;
    switch (id) {
        case RN_BUF_SYS_MAT_B: p = h->B; cnt = nx * nu; break;
        case RN_BUF_SYS_MAT_L: p = h->L; cnt = nu * nv; break;
        case RN_BUF_SYS_MAT_LHAT: p = h->Lhat; cnt = nu * nd; break;
        case RN_BUF_SYS_MAT_F:
        case RN_BUF_SYS_MAT_G: {
            if (!h->factored) return rn::fail(h, RN_ERR_STATE, "dense sysF/sysG requested before rn_factor_step");
            RN_CUDA(h, cudaSetDevice(h->device));
            RN_CHECK(rn::materialise_dense_sys(h));
            if (id == RN_BUF_SYS_MAT_F) { p = h->sysF_dense; cnt = n * 2 * nx * nx; } else { p = h->sysG_dense; cnt = n * nu * nu; }
            break;
        }
        case RN_BUF_SYS_XMIN: p = h->sxmin; cnt = n * nx; break;
        case RN_BUF_SYS_XMAX: p = h->sxmax; cnt = n * nx; break;
        case RN_BUF_SYS_XS: p = h->sxs; cnt = n * nx; break;
        case RN_BUF_SYS_XS_UPPER: p = h->sxs_upper; cnt = n * nx; break;
        case RN_BUF_SYS_UMIN: p = h->sumin; cnt = n * nu; break;
        case RN_BUF_SYS_UMAX: p = h->sumax; cnt = n * nu; break;
        case RN_BUF_MAT_PHI: p = h->Phi; cnt = n * nv * 2 * nx; break;
        case RN_BUF_MAT_PSI: p = h->Psi; cnt = n * nv * nu; break;
        case RN_BUF_MAT_THETA: p = h->Theta; cnt = (size_t)h->n_omega * nv * nx; break;
        case RN_BUF_MAT_OMEGA: p = h->Omega; cnt = (size_t)h->n_omega * nv * nv; break;
        case RN_BUF_MAT_D: p = h->D; cnt = n * nv * 2 * nx; break;
        case RN_BUF_MAT_F: p = h->F; cnt = n * nv * nu; break;
        case RN_BUF_MAT_G: p = h->G; cnt = nv * nx; break;
        case RN_BUF_MAT_SIGMA: p = h->sigma; cnt = n * nv; break;
        case RN_BUF_MAT_WV: p = h->Wv; cnt = nu * nv; break;
        case RN_BUF_DIAG: p = h->diag; cnt = n * (2 * nx + nu); break;
        case RN_BUF_VEC_E: p = h->e; cnt = n * nx; break;
        case RN_BUF_VEC_UHAT: p = h->uhat; cnt = n * nu; break;
        case RN_BUF_VEC_ALPHA: p = h->alpha; cnt = n * nu; break;
        case RN_BUF_VEC_BETA: p = h->beta; cnt = n * nv; break;
        case RN_BUF_VEC_CURRENT_STATE: p = h->xcur; cnt = nx; break;
        case RN_BUF_VEC_PREV_CONTROL: p = h->uprev; cnt = nu; break;
        case RN_BUF_VEC_PREV_UHAT: p = h->uhat_prev; cnt = nu; break;
        case RN_BUF_VEC_PREV_DEMAND: p = h->dprev; cnt = nd; break;
        case RN_BUF_VEC_X: p = h->X; cnt = n * nx; break;
        case RN_BUF_VEC_U: p = h->U; cnt = n * nu; break;
        case RN_BUF_VEC_V: p = h->V; cnt = n * nv; break;
        case RN_BUF_VEC_XI: p = h->xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_PSI: p = h->psi; cnt = n * nu; break;
        case RN_BUF_VEC_ACCEL_XI: p = h->acc_xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_ACCEL_PSI: p = h->acc_psi; cnt = n * nu; break;
        case RN_BUF_VEC_PRIMAL_XI: p = h->pri_xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_PRIMAL_PSI: p = h->pri_psi; cnt = n * nu; break;
        case RN_BUF_VEC_DUAL_XI: p = h->dual_xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_DUAL_PSI: p = h->dual_psi; cnt = n * nu; break;
        case RN_BUF_VEC_UPDATE_XI: p = h->upd_xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_UPDATE_PSI: p = h->upd_psi; cnt = n * nu; break;
        case RN_BUF_VEC_RESIDUAL_XI: p = h->res_xi; cnt = n * 2 * nx; break;
        case RN_BUF_VEC_RESIDUAL_PSI: p = h->res_psi; cnt = n * nu; break;
        case RN_BUF_CONTROL_ACTION: p = h->control_action; cnt = nu; break;
        case RN_BUF_STATE_UPDATE: p = h->state_update; cnt = nx; break;
        case RN_BUF_VEC_ZETA: p = h->zeta; cnt = n * nu; break;
        default: return rn::fail(h, RN_ERR_INVALID, "rn_buffer: unknown id %d", (int)id);
    }
    *dev_ptr = p;
    *bytes = cnt * sizeof(float);
    return RN_OK;
}

rn_status rn_read_buffer(rn_handle *hh, rn_buffer_id id, float *host, size_t count) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    void *p; size_t bytes;
    if (!h || !host) return RN_ERR_INVALID;
    RN_CHECK(rn_buffer(hh, id, &p, &bytes));
    if (count * sizeof(float) > bytes) return rn::fail(h, RN_ERR_INVALID, "rn_read_buffer: %zu floats > buffer (%zu)", count, bytes / 4);
    RN_CUDA(h, cudaMemcpyAsync(host, p, count * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_write_buffer(rn_handle *hh, rn_buffer_id id, const float *host, size_t count) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    void *p; size_t bytes;
    if (!h || !host) return RN_ERR_INVALID;
    RN_CHECK(rn_buffer(hh, id, &p, &bytes));
    if (count * sizeof(float) > bytes) return rn::fail(h, RN_ERR_INVALID, "rn_write_buffer: %zu floats > buffer (%zu)", count, bytes / 4);
    RN_CUDA(h, cudaMemcpyAsync(p, host, count * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RN_CUDA(h, cudaStreamSynchronize(h->stream));
    return RN_OK;
}

rn_status rn_profile_stream(rn_handle *hh, int reps, float *mean_ms) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || reps <= 0 || !mean_ms) return RN_ERR_INVALID;
    if (!h->factored) return rn::fail(h, RN_ERR_STATE, "rn_profile_stream before rn_factor_step");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::profile_stream(h, reps, mean_ms);
}

rn_status rn_profile_kernels(rn_handle *hh, int iterations, float *ms_out) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || iterations <= 0 || !ms_out) return RN_ERR_INVALID;
    if (!h->factored || !h->eliminated) return rn::fail(h, RN_ERR_STATE, "rn_profile_kernels before factor step / eliminate");
    RN_CUDA(h, cudaSetDevice(h->device));
    return rn::profile_kernels(h, iterations, ms_out);
}

rn_status rn_phase_times(rn_handle *hh, double *out) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !out) return RN_ERR_INVALID;
    for (int k = 0; k < 32; k++) out[k] = h->last_phase_iters > 0 ? (double)h->last_phase_ns[k] / h->last_phase_iters : 0.0;
    return RN_OK;
}

rn_status rn_cta_times(rn_handle *hh, double *out, int cap_ctas, int *n_ctas) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !out || !n_ctas) return RN_ERR_INVALID;
    const int n = (int)(h->last_cta_ns.size() / 2);
    *n_ctas = n;
    for (int k = 0; k < n && k < cap_ctas; k++) {
        out[2 * k] = h->last_phase_iters > 0 ? (double)h->last_cta_ns[2 * k] / h->last_phase_iters : 0.0;
        out[2 * k + 1] = (double)h->last_cta_ns[2 * k + 1];
    }
    return RN_OK;
}

}  // extern "C"
