// rn_internal.h -- handle layout and helpers shared by the CUDA translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rapidnet_b200.h"

namespace rn {

// Everything the kernels need to know about the tree, device-resident (0-based ids).
struct DevTree {
    int *stages = nullptr;        // [nodes]
    int *parent = nullptr;        // [nodes]  -1 for the root
    int *child_first = nullptr;   // [nodes]  first child id (0-based), children are contiguous
    int *child_count = nullptr;   // [nodes]  0 for leaves
    int *omega_idx = nullptr;     // [nodes]  which distinct Omega/Theta the node uses (Engine.cu:210-221)
    float *prob = nullptr;        // [nodes]
    float *err_demand = nullptr;  // [nodes*nd]
    float *err_price = nullptr;   // [nodes*nu]
};

struct Handle {
    rn_dims d{};
    int device = 0;
    int sm_count = 0;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t cap_stream = nullptr;   // private stream used only to capture the iteration graph
    bool own_stream = false;
    std::string err;

    // host copies
    std::vector<int> h_stages, h_nps, h_cum, h_parent, h_child_first, h_child_count, h_omega_idx;
    std::vector<float> h_prob, h_B, h_Gd, h_E, h_Ed, h_xmin, h_xmax, h_xsafe, h_umin, h_umax, h_alpha1;
    std::vector<float> h_W, h_precond, h_L, h_Lhat;
    bool have_null_space = false;
    float pen_x = 0, pen_xs = 0, step = 0, w_econ = 1;
    int max_iter = 0;
    int fb_node = 0, fb_stage = 0;   // ScenarioTree::getFinalBranchNode/Stage
    int n_omega = 0;                 // distinct Omega/Theta
    int chain_stage = 0;             // first stage of the non-branching tail (N if none)
    bool demand_uncertainty = true, price_uncertainty = true;
    bool factored = false, state_set = false, eliminated = false;
    int sweep_mode = RN_SWEEP_PERSISTENT, factor_mode = RN_FACTORS_FULL;

    DevTree t;
    // network / config constants on the device
    float *B = nullptr, *Gd = nullptr, *L = nullptr, *Lhat = nullptr, *W = nullptr, *Wv = nullptr;
    float *Rbar = nullptr, *G = nullptr /* Bbar' nv*nx */;
    float *OmegaBar = nullptr /* Rbar^-1 */, *ThetaBar = nullptr /* -1/2 Rbar^-1 Bbar' */, *PsiBar = nullptr /* -1/2 Rbar^-1 L' */, *Lt = nullptr /* L' nv*nu */;
    float *precond = nullptr, *alpha1 = nullptr;
    float *xmin = nullptr, *xmax = nullptr, *xsafe = nullptr, *umin = nullptr, *umax = nullptr;  // unscaled
    // factor-step outputs (reference layouts, + 16 B slack at the end of each streamed array)
    float *Phi = nullptr, *Psi = nullptr, *D = nullptr, *F = nullptr, *Omega = nullptr, *Theta = nullptr;
    float *factor_slab = nullptr;    // D | F | Phi | Psi live in one allocation
    size_t factor_df_bytes = 0, factor_full_bytes = 0;
    float *diag = nullptr;           // [nodes][2nx+nu] : s_x | s_xs | s_u
    float *sxmin = nullptr, *sxmax = nullptr, *sxs = nullptr, *sxs_upper = nullptr, *sumin = nullptr, *sumax = nullptr;
    float *sysF_dense = nullptr, *sysG_dense = nullptr;   // lazily materialised for the getters
    // per solve
    float *xcur = nullptr, *uprev = nullptr, *dprev = nullptr, *uhat_prev = nullptr;
    float *dhat = nullptr, *alphahat = nullptr;   // staged forecasts N*nd, N*nu
    float *e = nullptr, *uhat = nullptr, *alpha = nullptr, *beta = nullptr, *zeta = nullptr;
    // APG state
    float *X = nullptr, *U = nullptr, *V = nullptr, *sigma = nullptr;
    float *xi = nullptr, *psi = nullptr, *upd_xi = nullptr, *upd_psi = nullptr, *acc_xi = nullptr, *acc_psi = nullptr;
    float *pri_xi = nullptr, *pri_psi = nullptr, *dual_xi = nullptr, *dual_psi = nullptr, *res_xi = nullptr, *res_psi = nullptr;
    float *control_action = nullptr, *state_update = nullptr;
    // the ten dual/primal vectors live in one slab (one memset per cold start).  xi/psi/upd_* above are ROLE
    // pointers (y_{k-1}, y_k) into the two physical iterate buffers yA / yB, which swap roles every iteration.
    float *apg_slab = nullptr;
    size_t apg_slab_bytes = 0;
    float *yA_xi = nullptr, *yA_psi = nullptr, *yB_xi = nullptr, *yB_psi = nullptr;
    float *wA_xi = nullptr, *wA_psi = nullptr, *wB_xi = nullptr, *wB_psi = nullptr;   // w double buffer (persistent kernel)
    int *cum_dev = nullptr;          // nodes_per_stage_cumul on the device
    // sweep scratch
    float *a = nullptr, *b = nullptr, *c = nullptr;   // hoisted per-node products: nodes*nv, nodes*nv, nodes*nx
    float *q = nullptr, *r = nullptr;                 // nodes*nx, nodes*nv
    // reductions / loop control
    double *dist_part = nullptr;     // [2 * dist_slots] partial sums of squares (x-box, x-safe)
    int dist_slots = 0;
    float *scal = nullptr;           // [8] : d1, d2, scale1, scale2, flags...
    int *iter_dev = nullptr;         // device iteration counter
    unsigned int *done_ctr = nullptr;// last-block counters
    float *lambda_tab = nullptr;     // [max lambda entries]
    int lambda_cap = 0;
    int lambda_ready = 0;
    float *pinf = nullptr;           // [lambda_cap] primal infeasibility log
    float *pinf_part = nullptr;      // per-CTA candidates (abs, signed, idx) * 2 blocks
    int pinf_slots = 0;
    float *pinned = nullptr;         // pinned host staging
    size_t pinned_floats = 0;

    // persistent cooperative kernel
    float *part[4] = {nullptr, nullptr, nullptr, nullptr};   // D xi_w, F psi_w, Phi xi_w, Psi psi_w
    float *sweep_pack = nullptr;
    // subtree partition across GPUs (rn_dist_prepare / rn_dist_connect)
    int dist_world = 1, dist_rank = 0, dist_K_glob = 0, dist_chain_off = 0;
    std::vector<int> dist_head_lo, dist_head_hi;   // per crown node: global chain-index range of the heads below it
    void *xchg = nullptr;                          // this rank's exchange buffer (one cudaMalloc -> one IPC handle)
    void *xchg_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool xchg_opened[8] = {false, false, false, false, false, false, false, false};
    unsigned int xepoch = 0;
    unsigned int s_total = 0;                      // S rows every rank's table has received in all launches so far (the in-kernel counter's base)
    float *pinf4 = nullptr;
    float *bat_y = nullptr, *bat_y3 = nullptr, *bat_x = nullptr;   // RN_SWEEP_BATCHED scratch: GEMM outputs, X = -1/2 (sigma + G q_bar)
    float *head_q = nullptr, *head_r = nullptr;    // G q / r of this rank's chain heads, nv floats each (persistent kernel)
    int pinf4_cap = 0;
    float *cm_c = nullptr, *cm_lv = nullptr, *cm_beta = nullptr, *cm_uhat = nullptr, *cm_e = nullptr;   // chain-major arrays
    int *crown_rng = nullptr, *pos_dev = nullptr, *crown_path = nullptr;
    unsigned int *grid_bar = nullptr;
    unsigned long long *phase_ns = nullptr, *cta_ns = nullptr;
    std::vector<unsigned long long> last_cta_ns;
    bool persist_ready = false;
    bool warm_start = false, have_duals = false;   // rn_set_warm_start; a solve has left duals to start from
    bool pack_dirty = true;          // the sweep pack must be re-copied from the factor-step outputs before the next launch
    unsigned long long last_phase_ns[32] = {0};
    int last_phase_iters = 0;
    int persist_grid = 0;
    int grid_limit = 0;              // rn_set_grid_limit: most CTAs of the persistent kernel (0 = one per SM)

    cudaGraphExec_t iter_graph = nullptr;
    int graph_sweep = -1, graph_factor = -1;

    long long launches = 0;
    long long launches_per_iter = 0;
    size_t device_bytes = 0;
    float last_stream_ms = 0;
    std::vector<void *> allocs;
};

extern thread_local std::string g_create_error;

inline rn_status fail(Handle *h, rn_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define RN_CUDA(h, call)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return rn::fail((h), RN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                                 \
    } while (0)

#define RN_CHECK(expr)                       \
    do {                                     \
        rn_status s__ = (expr);              \
        if (s__ != RN_OK) return s__;        \
    } while (0)

template <typename T>
rn_status dev_alloc(Handle *h, T **p, size_t count, bool zero = true) {
    size_t bytes = (count ? count : 1) * sizeof(T) + 64;   // slack: streamed arrays are read in 16-B windows
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return fail(h, RN_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    if (zero) {
        e = cudaMemsetAsync(q, 0, bytes, h->stream);
        if (e != cudaSuccess) return fail(h, RN_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    h->allocs.push_back(q);
    h->device_bytes += bytes;
    *p = reinterpret_cast<T *>(q);
    return RN_OK;
}

template <typename T>
rn_status upload(Handle *h, T *dst, const T *src, size_t count) {
    RN_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return RN_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// implemented in rn_factor.cu / rn_affine.cu / rn_apg.cu
rn_status factor_step(Handle *h);
rn_status materialise_dense_sys(Handle *h);
rn_status update_state(Handle *h, const float *x, const float *u_prev, const float *d_prev);
rn_status eliminate_coupling(Handle *h, const float *d_hat, const float *alpha_hat);
rn_status fix_beta(Handle *h, int first, int count, const float *zeta_rows);
rn_status apg_init(Handle *h);
rn_status apg_step(Handle *h, rn_step_kind kind, float lambda);
rn_status apg_enqueue(Handle *h, int iterations);
rn_status apg_continue(Handle *h, int iterations, const float *lambda_host);
rn_status apg_warm(Handle *h, int iterations);
rn_status apg_release_graph(Handle *h);
rn_status profile_stream(Handle *h, int reps, float *mean_ms);
rn_status profile_kernels(Handle *h, int iterations, float *ms_out);
bool persistent_supported(const Handle *h);
rn_status persistent_prepare(Handle *h);
rn_status ensure_xchg(Handle *h);
size_t xchg_err_offset(const Handle *h);
rn_status dist_crown_beta(Handle *h, int pull);
rn_status persistent_launch(Handle *h, cudaStream_t st, int iters);
rn_status batched_prepare(Handle *h);
rn_status launch_sweeps_batched(Handle *h, cudaStream_t st, bool fuse_prox, int *n_launch, int *n_slots, cudaEvent_t mid);
bool use_batched(const Handle *h);
rn_status clamp_control(Handle *h);
rn_status move_forward(Handle *h);
void fill_lambda_table(std::vector<float> &tab, int iters);
double stream_bytes_per_iteration(const Handle *h);
double apg_bytes_per_iteration(const Handle *h);

}  // namespace rn
