/*
 * rapidnet_b200.h -- C ABI of the B200-native APG stochastic-MPC path.
 *
 * The reference (GPUEngineering/RapidNet) has no FFI: its operator surface is the C++
 * class API of one binary.  This header is the boundary underneath that surface -- what
 * a maintainer of the reference would bind Engine / SmpcController to (see
 * INTEGRATION.md).  Every entry point names the reference interface it replaces.
 *
 * Conventions: extern "C", plain pointers and sizes, int status (0 = ok; message via
 * rn_last_error), opaque handle, caller-owned HOST buffers in, library-owned DEVICE
 * buffers inside, one CUDA stream per handle, no hidden globals.  All matrices are flat
 * column-major fp32, all per-node vectors node-major, exactly as the reference lays them
 * out.  Tree node ids in rn_tree are the JSON's 1-based ids.
 *
 * There is no CPU fallback: every call that computes needs a CUDA device and fails with
 * RN_ERR_CUDA otherwise.
 */
#ifndef RAPIDNET_B200_H_
#define RAPIDNET_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rn_handle rn_handle;
typedef int rn_status;

enum {
    RN_OK = 0,
    RN_ERR_INVALID = 1,   /* bad argument / inconsistent dimensions */
    RN_ERR_CUDA = 2,      /* CUDA runtime / cuSOLVER failure (no device, launch error, ...) */
    RN_ERR_STATE = 3,     /* call order violated (e.g. solve before factor step) */
    RN_ERR_SINGULAR = 4,  /* p*Rbar not invertible (reference: Engine.cu:1334-1353 exits) */
    RN_ERR_NOMEM = 5
};

/* DwnNetwork / ScenarioTree / SmpcConfiguration dimensions (DwnNetwork.cuh:67-90,
 * ScenarioTree.cuh:64-100, SmpcConfiguration.cuh:60-80). */
typedef struct rn_dims {
    int nx, nu, nd, ne, nv;   /* tanks, controls, demands, mixing nodes, nv = nu - ne */
    int N, K, nodes;          /* horizon, scenarios, tree nodes */
    int n_nonleaf;            /* nNonLeafNodes */
    int n_children_tot;       /* nChildrenTot */
    int chain_stage_hint;     /* 0: whole tree (the reference's case).  > 0: this tree is one rank's part of a larger
                                 tree (crown + some of its scenario chains, rn_dist_*) whose non-branching tail starts at
                                 this stage; the Omega/Theta aliasing (Engine.cu:210-221) then follows the larger tree */
} rn_dims;

/* ScenarioTree getters (ScenarioTree.cuh:100-154); arrays as they appear in the JSON. */
typedef struct rn_tree {
    const int *stages;                 /* [nodes]   0-based stage of each node            */
    const int *nodes_per_stage;        /* [N+1]     trailing 0                             */
    const int *nodes_per_stage_cumul;  /* [N+2]                                            */
    const int *leaves;                 /* [K]       1-based                                */
    const int *children;               /* [n_children_tot] 1-based                         */
    const int *ancestor;               /* [nodes]   1-based parent id, root = 0            */
    const int *n_children;             /* [n_nonleaf]                                      */
    const int *n_children_cumul;       /* [nodes]                                          */
    const float *prob;                 /* [nodes]   path probability                       */
    const float *err_demand;           /* [nodes*nd]                                       */
    const float *err_price;            /* [nodes*nu]                                       */
} rn_tree;

/* DwnNetwork getters (DwnNetwork.cuh:91-143). matA is never used by the reference. */
typedef struct rn_network {
    const float *B;      /* nx*nu */
    const float *Gd;     /* nx*nd */
    const float *E;      /* ne*nu */
    const float *Ed;     /* ne*nd */
    const float *xmin, *xmax, *xsafe;   /* nx */
    const float *umin, *umax;           /* nu */
    const float *alpha1;                /* nu  (costAlpha1) */
} rn_network;

/* SmpcConfiguration getters (SmpcConfiguration.cuh:81-181). */
typedef struct rn_config {
    const float *costW;     /* nu*nu                                        */
    const float *precond;   /* N*(nu+2nx), per stage [u | x | xsafe]        */
    float penalty_x;        /* getPenaltyState()                            */
    float penalty_xs;       /* getPenaltySafety()                           */
    float step_size;        /* getStepSize()                                */
    float weight_economical;/* getWeightEconomical() (hard-coded 1)         */
    int max_iterations;     /* getMaxIterations()                           */
} rn_config;

/* How the solve step walks the tree (all give the same iterates up to fp32 rounding). */
typedef enum rn_sweep_mode {
    RN_SWEEP_PER_STAGE = 0,  /* one launch per stage, batched GEMV over the stage's nodes: the
                                literal shape of SmpcController::solveStep (:593-741)           */
    RN_SWEEP_CHAIN = 1,      /* node-parallel factor stream, then one CTA per scenario chain below
                                the last branching stage, per-stage launches above (CUDA graph)   */
    RN_SWEEP_PERSISTENT = 2, /* default: the whole APG loop in ONE persistent cooperative kernel, grid
                                barrier per crown stage, chains as scans + GEMMs across their stages;
                                falls back to RN_SWEEP_BATCHED when the problem does not fit it     */
    RN_SWEEP_BATCHED = 3     /* node-parallel factor stream, then the sweeps as four GEMMs ACROSS ALL NODES
                                against the shared matrices G, OmegaBar, L, B (every Omega_i is OmegaBar / p_i)
                                with element-wise scans over the stages in between (CUDA graph); any dimensions */
} rn_sweep_mode;

/* Which per-node factor matrices the stream kernel reads (DESIGN.md "formulations"). */
typedef enum rn_factor_mode {
    RN_FACTORS_FULL = 0,     /* Phi, Psi, D, F  (what solveStep reads; Tier-A bytes)            */
    RN_FACTORS_DF = 1,       /* D, F only; v = -1/2 Omega r (exact identity, half the bytes)    */
    RN_FACTORS_SHARED = 2    /* no per-node matrix is read ("Tier B"): D_i = G sysF_i',
                                F_i = L' sysG_i' (Engine.cu:720-728) with sysF, sysG diagonal, so
                                D xi = G (sysF' xi) and F psi = L' (s_u o psi) come from the two SHARED
                                matrices held in shared memory; v = -1/2 Omega r as in RN_FACTORS_DF.
                                Persistent sweep only; rounding differs from the streamed products  */
} rn_factor_mode;

/* Single APG sub-steps, for the golden-vector tests (TestSmpcController.cu:114-398 calls
 * the protected methods of SmpcController one by one). */
typedef enum rn_step_kind {
    RN_STEP_EXTRAPOLATE = 0,   /* SmpcController::dualExtrapolationStep(lambda)  :535-557 */
    RN_STEP_SOLVE = 1,         /* SmpcController::solveStep()                    :563-755 */
    RN_STEP_PROX = 2,          /* SmpcController::proximalFunG()                 :759-835 */
    RN_STEP_RESIDUAL = 3,      /* SmpcController::computeFixedPointResidual()    :839-850 */
    RN_STEP_DUAL_UPDATE = 4    /* SmpcController::dualUpdate()                   :854-864 */
} rn_step_kind;

/* Device buffers in the reference's packed layouts (Engine.cuh:104-318 getters and the
 * protected members of SmpcController.cuh:262-462 that TestSmpcController pokes). */
typedef enum rn_buffer_id {
    /* Engine constants */
    RN_BUF_SYS_MAT_B = 0,      /* getSysMatB      nx*nu                    */
    RN_BUF_SYS_MAT_L,          /* getSysMatL      nu*nv                    */
    RN_BUF_SYS_MAT_LHAT,       /* getSysMatLhat   nu*nd                    */
    RN_BUF_SYS_MAT_F,          /* getSysMatF      nodes*(2nx*nx) dense, materialised on request */
    RN_BUF_SYS_MAT_G,          /* getSysMatG      nodes*(nu*nu)  dense, materialised on request */
    RN_BUF_SYS_XMIN,           /* getSysXmin      nodes*nx  (preconditioned)                     */
    RN_BUF_SYS_XMAX,
    RN_BUF_SYS_XS,
    RN_BUF_SYS_XS_UPPER,       /* getSysXsUpper   nodes*nx  0x7F7F7F7F                           */
    RN_BUF_SYS_UMIN,
    RN_BUF_SYS_UMAX,
    RN_BUF_MAT_PHI,            /* getMatPhi       nodes*(nv*2nx)           */
    RN_BUF_MAT_PSI,            /* getMatPsi       nodes*(nv*nu)            */
    RN_BUF_MAT_THETA,          /* getMatTheta     fb*(nv*nx)               */
    RN_BUF_MAT_OMEGA,          /* getMatOmega     fb*(nv*nv)               */
    RN_BUF_MAT_D,              /* getMatD         nodes*(nv*2nx)           */
    RN_BUF_MAT_F,              /* getMatF         nodes*(nv*nu)            */
    RN_BUF_MAT_G,              /* getMatG         nv*nx (the reference keeps K identical copies) */
    RN_BUF_MAT_SIGMA,          /* getMatSigma     nodes*nv                 */
    RN_BUF_MAT_WV,             /* devMatWv        nu*nv                    */
    RN_BUF_DIAG,               /* private: nodes*(2nx+nu) = diag(sysF)|diag(sysG) per node      */
    /* per-solve affine terms */
    RN_BUF_VEC_E,              /* getVecE         nodes*nx                 */
    RN_BUF_VEC_UHAT,           /* getVecUhat      nodes*nu                 */
    RN_BUF_VEC_ALPHA,          /* getPriceAlpha   nodes*nu                 */
    RN_BUF_VEC_BETA,           /* getVecBeta      nodes*nv                 */
    RN_BUF_VEC_CURRENT_STATE,  /* getVecCurrentState      nx               */
    RN_BUF_VEC_PREV_CONTROL,   /* getVecPreviousControl   nu               */
    RN_BUF_VEC_PREV_UHAT,      /* getVecPreviousUhat      nu               */
    RN_BUF_VEC_PREV_DEMAND,    /* getVecPreviousDemand    nd               */
    /* SmpcController state */
    RN_BUF_VEC_X,              /* devVecX         nodes*nx                 */
    RN_BUF_VEC_U,              /* devVecU         nodes*nu                 */
    RN_BUF_VEC_V,              /* devVecV         nodes*nv                 */
    RN_BUF_VEC_XI,             /* devVecXi        nodes*2nx   y_{k-1}      */
    RN_BUF_VEC_PSI,            /* devVecPsi       nodes*nu                 */
    RN_BUF_VEC_ACCEL_XI,       /* devVecAcceleratedXi    w                 */
    RN_BUF_VEC_ACCEL_PSI,
    RN_BUF_VEC_PRIMAL_XI,      /* devVecPrimalXi         Hx                */
    RN_BUF_VEC_PRIMAL_PSI,
    RN_BUF_VEC_DUAL_XI,        /* devVecDualXi           z                 */
    RN_BUF_VEC_DUAL_PSI,
    RN_BUF_VEC_UPDATE_XI,      /* devVecUpdateXi         y_k               */
    RN_BUF_VEC_UPDATE_PSI,
    RN_BUF_VEC_RESIDUAL_XI,    /* devVecFixedPointResidualXi               */
    RN_BUF_VEC_RESIDUAL_PSI,
    RN_BUF_CONTROL_ACTION,     /* devControlAction       nu                */
    RN_BUF_STATE_UPDATE,       /* devStateUpdate         nx                */
    RN_BUF_VEC_ZETA,           /* devVecZeta             nodes*nu   (calculateZeta, Utilities.cu:100-131) */
    RN_BUF_COUNT_
} rn_buffer_id;

typedef struct rn_info {
    int device, sm_count;
    int final_branch_node;      /* ScenarioTree::getFinalBranchNode()  */
    int final_branch_stage;     /* ScenarioTree::getFinalBranchStage() */
    int chain_first_stage;      /* first stage of the non-branching tail; N if the tree has none */
    int num_omega;              /* distinct Omega/Theta matrices */
    int sweep_mode, factor_mode;
    long long kernel_launches;  /* kernels launched by this handle so far */
    long long launches_per_iteration;
    size_t device_bytes;        /* library-owned device memory */
    size_t factor_bytes;        /* bytes of Phi,Psi,D,F */
    double stream_bytes_per_iteration;  /* algorithmic bytes the stream kernel moves per APG iteration */
    double apg_bytes_per_iteration;     /* SURVEY 8(d) Tier-A figure for this problem and factor mode   */
    float last_distance_x, last_distance_xs;   /* prox distances d1, d2 of the last iteration */
    float last_stream_ms;       /* rn_profile_stream result */
} rn_info;

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* Engine::Engine + SmpcController::SmpcController allocation (Engine.cu:126-380,
 * SmpcController.cu:117-232): copies the host arrays, allocates every device buffer once. */
rn_status rn_create(const rn_dims *dims, const rn_tree *tree, const rn_network *net,
                    const rn_config *cfg, int device, rn_handle **out);
/* Engine::~Engine / SmpcController::~SmpcController */
rn_status rn_destroy(rn_handle *h);
/* message of the last failing call on `h` (or of the last failing rn_create if h == NULL) */
const char *rn_last_error(const rn_handle *h);
/* run all work of this handle on an existing cudaStream_t (default: a stream the handle owns) */
rn_status rn_set_stream(rn_handle *h, void *cuda_stream);
rn_status rn_get_stream(rn_handle *h, void **cuda_stream);
rn_status rn_sync(rn_handle *h);
rn_status rn_set_modes(rn_handle *h, rn_sweep_mode sweep, rn_factor_mode factors);
rn_status rn_get_info(rn_handle *h, rn_info *info);

/* ---- Engine --------------------------------------------------------------------------------- */

/* Optional: use this null-space basis L (nu*nv) and particular solution Lhat (nu*nd) instead of
 * computing them (Engine::calculateMatLandMatLhat, Engine.cu:466-669, cuSOLVER Dgesvd).  Lets a
 * test feed the basis the golden vectors were produced with. */
rn_status rn_set_null_space(rn_handle *h, const float *L, const float *Lhat);
/* Engine::factorStep() (Engine.cu:671-774) incl. initialiseSystemDevice (:382-463) */
rn_status rn_factor_step(rn_handle *h);
/* Engine::updateStateControl(currentX, prevU, prevDemand) (Engine.cu:1300-1316) */
rn_status rn_update_state(rn_handle *h, const float *x, const float *u_prev, const float *d_prev);
/* Engine::eliminateInputDistubanceCoupling(nominalDemand[N*nd], nominalPrices[N*nu]) (Engine.cu:1147-1298) */
rn_status rn_eliminate_coupling(rn_handle *h, const float *d_hat, const float *alpha_hat);
/* Engine::setDemandUncertaintyFlag / setPriceUncertaintyFlag (Engine.cu:1135-1145) */
rn_status rn_set_uncertainty(rn_handle *h, int demand, int price);

/* ---- SmpcController -------------------------------------------------------------------------- */

/* SmpcController::initialiseAlgorithm (SmpcController.cu:420-450): zero all duals (cold start) */
rn_status rn_apg_init(rn_handle *h);
/* one protected step method (see rn_step_kind); lambda is used by RN_STEP_EXTRAPOLATE only */
rn_status rn_step(rn_handle *h, rn_step_kind kind, float lambda);
/* SmpcController::algorithmApg() (SmpcController.cu:1500-1525): cold start + `iterations` fused
 * iterations.  Asynchronous on the handle's stream unless a host output is requested:
 *   u0_host          nullable, nu floats   = devVecU[0:nu]
 *   primal_infs_host nullable, iterations floats = vecPrimalInfs (updatePrimalInfeasibity :1480-1496) */
rn_status rn_apg_solve(rn_handle *h, int iterations, float *u0_host, float *primal_infs_host);
/* rn_apg_solve's primal_infs_host is this rank's log when the tree is partitioned (rn_dist_prepare with world > 1);
 * rn_read_pinf_parts + the host merge give the global one.
 *
 * Continue from the duals in place instead of cold-starting (no counterpart in the reference, whose algorithmApg always
 * zeroes them, :1509; opt-in).  Persistent sweep only; asynchronous.
 *   warm_restart == 0: devVecUpdateXi/Psi is y_k, devVecXi/Psi is y_{k-1}; lambda_host (nullable, `iterations` floats)
 *                      replaces the theta recursion for these iterations.  With one iteration and the golden inputs of
 *                      TestSmpcController.cu:134-398 this runs extrapolate -> solve -> prox -> residual -> update through
 *                      the persistent kernel.
 *   warm_restart != 0: warm start of a closed-loop step: y_0 = y_{-1} = the duals the previous solve left, theta restarts
 *                      at 1 (lambda_host ignored). */
rn_status rn_apg_continue(rn_handle *h, int iterations, const float *lambda_host, int warm_restart);
/* Opt-in: from the second solve on, rn_apg_solve / rn_control_action start from the duals the previous solve left (theta
 * restarted) instead of zeroing them -- for receding-horizon loops, where consecutive problems differ little.  Off by default:
 * the reference cold-starts every solve (:420-450, :1509) and parity is defined on that.  Persistent sweep only; a factor step
 * or switching it off drops the stored duals. */
rn_status rn_set_warm_start(rn_handle *h, int on);
/* SmpcController::controlAction(real_t* u) (:1607-1625) when clamp == 0;
 * SmpcController::controlAction(fstream&) control vector (:1633-1667) when clamp != 0 (u0 clamped with
 * the node-0 preconditioned bounds).  Host buffers in, u0 (nu floats) out; blocking. */
rn_status rn_control_action(rn_handle *h, const float *x, const float *u_prev, const float *d_prev,
                            const float *d_hat, const float *alpha_hat, int iterations, int clamp,
                            float *u0_host);
/* plant update of SmpcController::moveForewardInTime (:1679-1717): x_next = x + B*u0_clamped
 * (reference drops the disturbance term).  Outputs nx and nu host floats. */
rn_status rn_move_forward(rn_handle *h, float *x_next_host, float *u_applied_host);

/* ---- one tree across several GPUs (no counterpart in the reference, which is single-GPU) ------------------------
 * Subtree partition: every rank (one process per GPU) holds the stages above the last branching stage (the "crown",
 * replicated) and a contiguous range of the scenario chains below it; its handle is created on that LOCAL tree.  Per
 * APG iteration the ranks exchange, inside the persistent kernel and over NVLink peer memory, the sums of q and r over
 * the chain heads below each node of the last crown stage (the near-root contributions of SmpcController::solveStep's
 * backward sweep, :661-672 -- what solveSumChildren leaves in the parent's slot) and the two squared prox distances
 * (:792, :810).  Call order: rn_create (local tree) -> rn_dist_prepare -> exchange the 64-byte
 * handles between the processes (e.g. torch.distributed.all_gather_object) -> rn_dist_connect -> factor step, solves. */
#define RN_IPC_HANDLE_BYTES 64
/* world <= 8 ranks; this rank owns global chains [chain_offset, chain_offset + K_local) of K_global; head_lo/head_hi
 * [number of crown nodes]: global chain-index range of the chain heads below each crown node (used to check that the
 * chains below every node of the last crown stage sit on ONE rank).  Writes the CUDA IPC handle of this rank's exchange
 * buffer to ipc_handle_out.  Per iteration a rank sends (nx + nv) floats per node of the last crown stage it owns. */
rn_status rn_dist_prepare(rn_handle *h, int world, int rank, int K_global, int chain_offset, const int *head_lo,
                          const int *head_hi, unsigned char *ipc_handle_out /*[RN_IPC_HANDLE_BYTES]*/);
/* handles: world x RN_IPC_HANDLE_BYTES, rank-major (the entry of this rank is ignored) */
rn_status rn_dist_connect(rn_handle *h, const unsigned char *handles);
/* The only per-SOLVE quantity that couples the ranks: zeta of a node just above the chain heads sums over ALL of
 * its children (calculateZeta, Utilities.cu:100-131), some of which live on other ranks.  After rn_eliminate_coupling
 * the host sums the ranks' contributions and hands the corrected zeta rows of nodes [first, first + count) back;
 * beta = 2 (W L)' zeta + p L' alpha (Engine.cu:1253-1261) is recomputed for them. */
rn_status rn_dist_fix_crown_beta(rn_handle *h, int first, int count, const float *zeta_rows /*[count*nu]*/);
/* The same coupling without the host in the data path.  The partition is aligned to the nodes just above the chain heads
 * (rn_dist_prepare rejects one that is not), so the rank that holds a node's chains has its exact zeta / beta row:
 *   pull == 0: store this rank's rows into every rank's staging table (peer memory); then synchronise the stream and
 *              pass a barrier across the ranks;
 *   pull != 0: copy the rows this rank does not own from its staging table into beta.
 * Asynchronous on the handle's stream. */
rn_status rn_dist_sync_crown_beta(rn_handle *h, int pull);
/* per iteration (|res|, res) at the arg-max of the xi block and of the psi block of THIS rank's nodes, 4 floats each:
 * the host merges them across ranks into vecPrimalInfs (updatePrimalInfeasibity, :1480-1496) */
rn_status rn_read_pinf_parts(rn_handle *h, int iterations, float *host /*[iterations*4]*/);
/* nonzero when a cross-GPU wait of the last solve timed out (a peer died): results are invalid */
rn_status rn_dist_error(rn_handle *h, int *timed_out);

/* ---- buffers ---------------------------------------------------------------------------------- */

/* borrowed device pointer + size in bytes of a buffer in the reference's packed layout */
rn_status rn_buffer(rn_handle *h, rn_buffer_id id, void **dev_ptr, size_t *bytes);
/* blocking convenience copies (count in floats, from the start of the buffer) */
rn_status rn_read_buffer(rn_handle *h, rn_buffer_id id, float *host, size_t count);
rn_status rn_write_buffer(rn_handle *h, rn_buffer_id id, const float *host, size_t count);

/* ---- measurement ------------------------------------------------------------------------------ */

/* launch the factor-stream kernel `reps` times back to back on the handle's stream (state is left
 * untouched: lambda = 0 makes the extrapolation the identity) and return the mean duration in ms */
rn_status rn_profile_stream(rn_handle *h, int reps, float *mean_ms);

/* kernel classes of one APG iteration, for rn_profile_kernels */
typedef enum rn_prof_class {
    RN_PROF_STREAM = 0,     /* factor-matrix stream (+ fused dual extrapolation)          */
    RN_PROF_BACKWARD = 1,   /* backward tree sweep                                        */
    RN_PROF_FORWARD = 2,    /* forward tree sweep + Hx + box projections                  */
    RN_PROF_FINALIZE = 3,   /* distance branch, residual, dual update, infeasibility log  */
    RN_PROF_COUNT_ = 4
} rn_prof_class;
/* one cold-started solve of `iterations` iterations launched kernel by kernel with CUDA events (on the
 * launching stream) around each kernel class; ms_out[RN_PROF_COUNT_] = mean ms per iteration per class.
 * Leaves the same state behind as rn_apg_solve(iterations). */
rn_status rn_profile_kernels(rn_handle *h, int iterations, float *ms_out);
/* fine-grained phase clock of the last rn_profile_kernels run in persistent mode: 32 accumulators (ns per
 * iteration, CTA 0's %globaltimer); index meaning in DESIGN.md / rapidnet_b200/cabi.py PHASE_NAMES */
rn_status rn_phase_times(rn_handle *h, double *ns_per_iteration /*[32]*/);
/* Load balance of the factor stream in the last rn_profile_kernels run: out[2k] = ns per iteration CTA k spent in phase S,
 * out[2k+1] = the SM it ran on.  n_ctas = size of the persistent grid. */
rn_status rn_cta_times(rn_handle *h, double *out /*[2*cap_ctas]*/, int cap_ctas, int *n_ctas);
/* Allocates now what the first solve would allocate lazily (the persistent kernel's private buffers), so that the first
 * solve -- in particular the first launch of a partitioned tree, whose in-kernel waits span the GPUs -- does not. */
rn_status rn_prepare(rn_handle *h);
/* Caps the grid of the persistent kernel at max_ctas CTAs (0 = one per SM, the default).  Independent SMPC instances
 * (closed-loop Monte-Carlo, BASELINE config[3]) shard with no collective; on ONE GPU several handles, each on its own
 * stream with a share of the SMs, solve side by side: a tree with K scenarios keeps only K CTAs busy in its sweeps.
 * The reference runs one controller at a time (src/main.cu:27-63). */
rn_status rn_set_grid_limit(rn_handle *h, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* RAPIDNET_B200_H_ */
