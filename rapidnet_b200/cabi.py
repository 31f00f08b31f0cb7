"""ctypes binding of include/rapidnet_b200.h (the C ABI of the CUDA library).

Plumbing only: loads rapidnet_b200/librapidnet_b200.so (built in-tree by tools/build_lib.sh or
__graft_entry__.build()) and fails loudly if it is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAPIDNET_B200_LIB") or os.path.join(_HERE, "librapidnet_b200.so")   # override: A/B builds

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)

RN_OK = 0
STATUS_NAMES = {0: "RN_OK", 1: "RN_ERR_INVALID", 2: "RN_ERR_CUDA", 3: "RN_ERR_STATE", 4: "RN_ERR_SINGULAR",
                5: "RN_ERR_NOMEM"}

SWEEP_PER_STAGE, SWEEP_CHAIN, SWEEP_PERSISTENT, SWEEP_BATCHED = 0, 1, 2, 3
FACTORS_FULL, FACTORS_DF, FACTORS_SHARED = 0, 1, 2
STEP_EXTRAPOLATE, STEP_SOLVE, STEP_PROX, STEP_RESIDUAL, STEP_DUAL_UPDATE = range(5)

BUFFER_IDS = [
    "SYS_MAT_B", "SYS_MAT_L", "SYS_MAT_LHAT", "SYS_MAT_F", "SYS_MAT_G", "SYS_XMIN", "SYS_XMAX", "SYS_XS",
    "SYS_XS_UPPER", "SYS_UMIN", "SYS_UMAX", "MAT_PHI", "MAT_PSI", "MAT_THETA", "MAT_OMEGA", "MAT_D", "MAT_F",
    "MAT_G", "MAT_SIGMA", "MAT_WV", "DIAG", "VEC_E", "VEC_UHAT", "VEC_ALPHA", "VEC_BETA", "VEC_CURRENT_STATE",
    "VEC_PREV_CONTROL", "VEC_PREV_UHAT", "VEC_PREV_DEMAND", "VEC_X", "VEC_U", "VEC_V", "VEC_XI", "VEC_PSI",
    "VEC_ACCEL_XI", "VEC_ACCEL_PSI", "VEC_PRIMAL_XI", "VEC_PRIMAL_PSI", "VEC_DUAL_XI", "VEC_DUAL_PSI",
    "VEC_UPDATE_XI", "VEC_UPDATE_PSI", "VEC_RESIDUAL_XI", "VEC_RESIDUAL_PSI", "CONTROL_ACTION", "STATE_UPDATE", "VEC_ZETA",
]
BUF = {name: i for i, name in enumerate(BUFFER_IDS)}

# every symbol include/rapidnet_b200.h declares
EXPORTS = [
    "rn_create", "rn_destroy", "rn_last_error", "rn_set_stream", "rn_get_stream", "rn_sync", "rn_set_modes",
    "rn_get_info", "rn_set_null_space", "rn_factor_step", "rn_update_state", "rn_eliminate_coupling",
    "rn_set_uncertainty", "rn_apg_init", "rn_step", "rn_apg_solve", "rn_control_action", "rn_move_forward",
    "rn_buffer", "rn_read_buffer", "rn_write_buffer", "rn_profile_stream", "rn_profile_kernels", "rn_phase_times", "rn_cta_times",
    "rn_dist_prepare", "rn_dist_connect", "rn_dist_fix_crown_beta", "rn_read_pinf_parts", "rn_dist_error",
    "rn_set_grid_limit", "rn_apg_continue", "rn_dist_sync_crown_beta", "rn_prepare", "rn_set_warm_start",
]


class RnDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "nu", "nd", "ne", "nv", "N", "K", "nodes", "n_nonleaf", "n_children_tot",
                                       "chain_stage_hint")]


class RnTree(C.Structure):
    _fields_ = [(n, IP) for n in ("stages", "nodes_per_stage", "nodes_per_stage_cumul", "leaves", "children",
                                  "ancestor", "n_children", "n_children_cumul")] + \
               [(n, FP) for n in ("prob", "err_demand", "err_price")]


class RnNetwork(C.Structure):
    _fields_ = [(n, FP) for n in ("B", "Gd", "E", "Ed", "xmin", "xmax", "xsafe", "umin", "umax", "alpha1")]


class RnConfig(C.Structure):
    _fields_ = [("costW", FP), ("precond", FP), ("penalty_x", C.c_float), ("penalty_xs", C.c_float),
                ("step_size", C.c_float), ("weight_economical", C.c_float), ("max_iterations", C.c_int)]


class RnInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("final_branch_node", C.c_int),
                ("final_branch_stage", C.c_int), ("chain_first_stage", C.c_int), ("num_omega", C.c_int),
                ("sweep_mode", C.c_int), ("factor_mode", C.c_int), ("kernel_launches", C.c_longlong),
                ("launches_per_iteration", C.c_longlong), ("device_bytes", C.c_size_t),
                ("factor_bytes", C.c_size_t), ("stream_bytes_per_iteration", C.c_double),
                ("apg_bytes_per_iteration", C.c_double), ("last_distance_x", C.c_float),
                ("last_distance_xs", C.c_float), ("last_stream_ms", C.c_float)]


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with tools/build_lib.sh or "
                           f"`python -c 'import __graft_entry__ as g; g.build()'` -- rapidnet_b200 has no CPU path")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.rn_create.argtypes = [C.POINTER(RnDims), C.POINTER(RnTree), C.POINTER(RnNetwork), C.POINTER(RnConfig),
                              C.c_int, C.POINTER(H)]
    lib.rn_destroy.argtypes = [H]
    lib.rn_last_error.argtypes = [H]
    lib.rn_last_error.restype = C.c_char_p
    lib.rn_set_stream.argtypes = [H, C.c_void_p]
    lib.rn_get_stream.argtypes = [H, C.POINTER(C.c_void_p)]
    lib.rn_sync.argtypes = [H]
    lib.rn_set_modes.argtypes = [H, C.c_int, C.c_int]
    lib.rn_get_info.argtypes = [H, C.POINTER(RnInfo)]
    lib.rn_set_null_space.argtypes = [H, FP, FP]
    lib.rn_factor_step.argtypes = [H]
    lib.rn_update_state.argtypes = [H, FP, FP, FP]
    lib.rn_eliminate_coupling.argtypes = [H, FP, FP]
    lib.rn_set_uncertainty.argtypes = [H, C.c_int, C.c_int]
    lib.rn_apg_init.argtypes = [H]
    lib.rn_step.argtypes = [H, C.c_int, C.c_float]
    lib.rn_apg_solve.argtypes = [H, C.c_int, FP, FP]
    lib.rn_control_action.argtypes = [H, FP, FP, FP, FP, FP, C.c_int, C.c_int, FP]
    lib.rn_apg_continue.argtypes = [H, C.c_int, FP, C.c_int]
    lib.rn_move_forward.argtypes = [H, FP, FP]
    lib.rn_buffer.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.rn_read_buffer.argtypes = [H, C.c_int, FP, C.c_size_t]
    lib.rn_write_buffer.argtypes = [H, C.c_int, FP, C.c_size_t]
    lib.rn_profile_stream.argtypes = [H, C.c_int, FP]
    lib.rn_profile_kernels.argtypes = [H, C.c_int, FP]
    lib.rn_phase_times.argtypes = [H, C.POINTER(C.c_double)]
    lib.rn_cta_times.argtypes = [H, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    lib.rn_dist_prepare.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, IP, IP, C.c_char_p]
    lib.rn_dist_connect.argtypes = [H, C.c_char_p]
    lib.rn_dist_fix_crown_beta.argtypes = [H, C.c_int, C.c_int, FP]
    lib.rn_read_pinf_parts.argtypes = [H, C.c_int, FP]
    lib.rn_dist_error.argtypes = [H, IP]
    lib.rn_dist_sync_crown_beta.argtypes = [H, C.c_int]
    lib.rn_prepare.argtypes = [H]
    lib.rn_set_warm_start.argtypes = [H, C.c_int]
    for name in EXPORTS:
        if name != "rn_last_error":
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(FP)


def _ip(a):
    return a.ctypes.data_as(IP)


class RapidNetError(RuntimeError):
    pass


class Solver:
    """One SMPC problem on one GPU: Engine + SmpcController state behind the C ABI.

    Method names follow the reference (Engine::factorStep, ::updateStateControl,
    ::eliminateInputDistubanceCoupling; SmpcController::algorithmApg, ::controlAction)."""

    def __init__(self, problem, device: int = 0, chain_stage_hint: int = 0):
        lib = load()
        n, t, c = problem.network, problem.tree, problem.config
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.problem = problem
        self._keep = dict(
            stages=i32(t.stages), nps=i32(t.nodes_per_stage), cum=i32(t.nodes_per_stage_cumul), leaves=i32(t.leaves),
            children=i32(t.children), ancestor=i32(t.ancestor), nch=i32(t.n_children), ncc=i32(t.n_children_cumul),
            prob=f32(t.prob), ed=f32(t.err_demand), ep=f32(t.err_price), B=f32(n.B), Gd=f32(n.Gd), E=f32(n.E),
            Ed=f32(n.Ed), xmin=f32(n.xmin), xmax=f32(n.xmax), xsafe=f32(n.xsafe), umin=f32(n.umin), umax=f32(n.umax),
            a1=f32(n.alpha1), W=f32(c.costW), pc=f32(c.precond))
        k = self._keep
        dims = RnDims(n.nx, n.nu, n.nd, n.ne, c.nv, t.N, t.K, t.nodes, t.n_nonleaf, t.n_children_tot, int(chain_stage_hint))
        tree = RnTree(_ip(k["stages"]), _ip(k["nps"]), _ip(k["cum"]), _ip(k["leaves"]), _ip(k["children"]),
                      _ip(k["ancestor"]), _ip(k["nch"]), _ip(k["ncc"]), _fp(k["prob"]), _fp(k["ed"]), _fp(k["ep"]))
        net = RnNetwork(_fp(k["B"]), _fp(k["Gd"]), _fp(k["E"]), _fp(k["Ed"]), _fp(k["xmin"]), _fp(k["xmax"]),
                        _fp(k["xsafe"]), _fp(k["umin"]), _fp(k["umax"]), _fp(k["a1"]))
        cfg = RnConfig(_fp(k["W"]), _fp(k["pc"]), float(c.penalty_x), float(c.penalty_xs), float(c.step_size), 1.0,
                       int(c.max_iter))
        self.dims = dims
        self.h = C.c_void_p()
        rc = lib.rn_create(C.byref(dims), C.byref(tree), C.byref(net), C.byref(cfg), int(device), C.byref(self.h))
        if rc != RN_OK:
            msg = lib.rn_last_error(None)
            self.h = None
            raise RapidNetError(f"rn_create: {STATUS_NAMES.get(rc, rc)}: {msg.decode() if msg else ''}")

    # -- plumbing --
    def _check(self, rc, what):
        if rc != RN_OK:
            msg = load().rn_last_error(self.h)
            raise RapidNetError(f"{what}: {STATUS_NAMES.get(rc, rc)}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None):
            load().rn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._check(load().rn_set_stream(self.h, C.c_void_p(cuda_stream)), "rn_set_stream")

    def sync(self):
        self._check(load().rn_sync(self.h), "rn_sync")

    def set_modes(self, sweep=SWEEP_PERSISTENT, factors=FACTORS_FULL):
        self._check(load().rn_set_modes(self.h, sweep, factors), "rn_set_modes")

    def set_grid_limit(self, max_ctas: int):
        """Cap the persistent kernel's grid (0 = one CTA per SM) so that several handles can solve side by side."""
        self._check(load().rn_set_grid_limit(self.h, int(max_ctas)), "rn_set_grid_limit")

    def info(self) -> RnInfo:
        info = RnInfo()
        self._check(load().rn_get_info(self.h, C.byref(info)), "rn_get_info")
        return info

    # -- Engine --
    def set_null_space(self, L, Lhat):
        L = np.ascontiguousarray(L, dtype=np.float32)
        Lhat = np.ascontiguousarray(Lhat, dtype=np.float32)
        self._check(load().rn_set_null_space(self.h, _fp(L), _fp(Lhat)), "rn_set_null_space")

    def factor_step(self):
        self._check(load().rn_factor_step(self.h), "rn_factor_step")

    def update_state(self, x=None, u_prev=None, d_prev=None):
        c = self.problem.config
        x = np.ascontiguousarray(c.current_x if x is None else x, dtype=np.float32)
        u = np.ascontiguousarray(c.prev_u if u_prev is None else u_prev, dtype=np.float32)
        d = np.ascontiguousarray(c.prev_demand if d_prev is None else d_prev, dtype=np.float32)
        self._check(load().rn_update_state(self.h, _fp(x), _fp(u), _fp(d)), "rn_update_state")

    def eliminate_coupling(self, d_hat, alpha_hat):
        dh = np.ascontiguousarray(d_hat, dtype=np.float32)
        ah = np.ascontiguousarray(alpha_hat, dtype=np.float32)
        self._check(load().rn_eliminate_coupling(self.h, _fp(dh), _fp(ah)), "rn_eliminate_coupling")

    def set_uncertainty(self, demand=True, price=True):
        self._check(load().rn_set_uncertainty(self.h, int(demand), int(price)), "rn_set_uncertainty")

    # -- SmpcController --
    def apg_init(self):
        self._check(load().rn_apg_init(self.h), "rn_apg_init")

    def step(self, kind: int, lam: float = 0.0):
        self._check(load().rn_step(self.h, kind, float(lam)), "rn_step")

    def apg_solve(self, iterations: int, want_u0=True, want_infs=False):
        u0 = np.zeros(self.dims.nu, dtype=np.float32) if want_u0 else None
        infs = np.zeros(max(iterations, 1), dtype=np.float32) if want_infs else None
        self._check(load().rn_apg_solve(self.h, int(iterations), _fp(u0) if want_u0 else None,
                                        _fp(infs) if want_infs else None), "rn_apg_solve")
        return u0, (infs[:iterations] if want_infs else None)

    def set_warm_start(self, on: bool = True):
        """opt-in: later solves start from the duals of the previous one (the reference cold-starts)"""
        self._check(load().rn_set_warm_start(self.h, int(bool(on))), "rn_set_warm_start")

    def apg_continue(self, iterations: int, lambdas=None, warm_restart=False):
        """Iterations that continue from the duals in place (UPDATE = y_k, XI/PSI = y_{k-1}); `lambdas` replaces the theta
        recursion; warm_restart: y_0 = y_{-1} = the previous solve's duals, theta restarted.  Persistent sweep only."""
        lam = None if lambdas is None else np.ascontiguousarray(lambdas, dtype=np.float32)
        if lam is not None and lam.size < iterations:
            raise ValueError("apg_continue: one lambda per iteration")
        self._check(load().rn_apg_continue(self.h, int(iterations), _fp(lam) if lam is not None else None,
                                           int(bool(warm_restart))), "rn_apg_continue")
        self.sync()

    def control_action(self, x, u_prev, d_prev, d_hat, alpha_hat, iterations: int, clamp=False, out=None):
        args = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, u_prev, d_prev, d_hat, alpha_hat)]
        u0 = np.zeros(self.dims.nu, dtype=np.float32) if out is None else out
        self._check(load().rn_control_action(self.h, *[_fp(a) for a in args], int(iterations), int(bool(clamp)),
                                             _fp(u0)), "rn_control_action")
        return u0

    # -- one tree across several GPUs (rapidnet_b200/partition.py drives these) --
    def dist_prepare(self, world: int, rank: int, k_global: int, chain_offset: int, head_lo, head_hi) -> bytes:
        lo = np.ascontiguousarray(head_lo, dtype=np.int32)
        hi = np.ascontiguousarray(head_hi, dtype=np.int32)
        out = C.create_string_buffer(64)
        self._check(load().rn_dist_prepare(self.h, int(world), int(rank), int(k_global), int(chain_offset), _ip(lo), _ip(hi),
                                           out), "rn_dist_prepare")
        return out.raw

    def dist_connect(self, handles):
        blob = b"".join(handles)
        self._check(load().rn_dist_connect(self.h, blob), "rn_dist_connect")

    def dist_fix_crown_beta(self, first: int, zeta_rows):
        z = np.ascontiguousarray(zeta_rows, dtype=np.float32)
        self._check(load().rn_dist_fix_crown_beta(self.h, int(first), int(z.shape[0]), _fp(z)), "rn_dist_fix_crown_beta")

    def dist_sync_crown_beta(self, pull: bool):
        self._check(load().rn_dist_sync_crown_beta(self.h, int(bool(pull))), "rn_dist_sync_crown_beta")

    def prepare_persistent(self):
        """allocate what the first solve would allocate (so that the first timed / cross-GPU launch does not)"""
        self._check(load().rn_prepare(self.h), "rn_prepare")

    def pinf_parts(self, iterations: int) -> np.ndarray:
        out = np.zeros((max(iterations, 1), 4), dtype=np.float32)
        self._check(load().rn_read_pinf_parts(self.h, int(iterations), _fp(out)), "rn_read_pinf_parts")
        return out[:iterations]

    def dist_error(self) -> bool:
        flag = np.zeros(1, dtype=np.int32)
        self._check(load().rn_dist_error(self.h, _ip(flag)), "rn_dist_error")
        return bool(flag[0])

    def move_forward(self):
        x = np.zeros(self.dims.nx, dtype=np.float32)
        u = np.zeros(self.dims.nu, dtype=np.float32)
        self._check(load().rn_move_forward(self.h, _fp(x), _fp(u)), "rn_move_forward")
        return x, u

    # -- buffers --
    def buffer(self, name: str):
        p, nbytes = C.c_void_p(), C.c_size_t()
        self._check(load().rn_buffer(self.h, BUF[name], C.byref(p), C.byref(nbytes)), f"rn_buffer({name})")
        return p.value, nbytes.value

    def read(self, name: str, count=None) -> np.ndarray:
        _, nbytes = self.buffer(name)
        cnt = nbytes // 4 if count is None else int(count)
        out = np.zeros(cnt, dtype=np.float32)
        self._check(load().rn_read_buffer(self.h, BUF[name], _fp(out), cnt), f"rn_read_buffer({name})")
        return out

    def write(self, name: str, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        self._check(load().rn_write_buffer(self.h, BUF[name], _fp(a), a.size), f"rn_write_buffer({name})")

    def profile_stream(self, reps: int = 20) -> float:
        ms = C.c_float()
        self._check(load().rn_profile_stream(self.h, int(reps), C.byref(ms)), "rn_profile_stream")
        return float(ms.value)

    PROF_CLASSES = ("stream", "backward", "forward", "finalize")

    def profile_kernels(self, iterations: int = 50) -> dict:
        """mean ms per APG iteration of each kernel class (CUDA events on the launching stream)"""
        out = np.zeros(len(self.PROF_CLASSES), dtype=np.float32)
        self._check(load().rn_profile_kernels(self.h, int(iterations), _fp(out)), "rn_profile_kernels")
        return dict(zip(self.PROF_CLASSES, (float(v) for v in out)))

    PHASE_NAMES = {0: "S.stream", 1: "S.barrier", 2: "S.pinf", 3: "B.qscan", 4: "B.gemm_Gq", 5: "B.rscan", 6: "B.gemm_OT",
                   7: "B.vcombine", 8: "B.gemm_Lv", 9: "B.out", 10: "B.idle", 11: "C.sums", 12: "C.gemm_G", 13: "F.crown_path",
                   14: "F.crown_rest", 15: "F.columns", 16: "F.uscan", 17: "F.gemm_Bu", 18: "F.xscan", 19: "F.epilogue",
                   20: "C.barrier_in", 21: "C.idle", 22: "C.barrier_out", 23: "F.idle", 24: "cyc.gemv_wait_full",
                   25: "cyc.gemv_wait_w", 26: "cyc.gemv_wait_red", 27: "cyc.gemv_compute", 28: "F.dist", 29: "F.barrier",
                   30: "cyc.ew_wait_red", 31: "cyc.ew_prologue"}

    def cta_times(self):
        """(ns per iteration in phase S, SM id) of every CTA of the persistent grid in the last profile_kernels run"""
        out, n = (C.c_double * 2048)(), C.c_int(0)
        self._check(load().rn_cta_times(self.h, out, 1024, C.byref(n)), "rn_cta_times")
        return [(out[2 * k], int(out[2 * k + 1])) for k in range(n.value)]

    def phase_times(self) -> dict:
        """ns per iteration of each phase of the persistent kernel in the last profile_kernels run (CTA 0's clock)"""
        out = (C.c_double * 32)()
        self._check(load().rn_phase_times(self.h, out), "rn_phase_times")
        return {self.PHASE_NAMES.get(k, str(k)): out[k] for k in range(32) if out[k] > 0}
