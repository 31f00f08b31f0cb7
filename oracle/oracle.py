"""ctypes wrapper of oracle/rapidnet_oracle.c (test infrastructure, see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librapidnet_oracle.so")
_LIB_PATH64 = os.path.join(_HERE, "_build", "librapidnet_oracle64.so")

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rapidnet_oracle.c")
    if force or any(not os.path.exists(q) or os.path.getmtime(q) < os.path.getmtime(src) for q in (_LIB_PATH, _LIB_PATH64)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Lib:
    """ctypes view of one build of rapidnet_oracle.c: fp32 (the oracle proper) or fp64 (noise-floor yardstick)."""

    def __init__(self, path, f64):
        self.np = np.float64 if f64 else np.float32
        self.cf = C.c_double if f64 else C.c_float
        self.RP = C.POINTER(self.cf)
        RP, cf = self.RP, self.cf
        l = self.l = C.CDLL(path)
        l.orc_create.restype = C.c_void_p
        l.orc_create.argtypes = [C.c_int] * 9 + [IP] * 6 + [RP] * 3 + [RP] * 10 + [RP, RP, cf, cf, cf]
        l.orc_destroy.argtypes = [C.c_void_p]
        l.orc_set_L.argtypes = [C.c_void_p, RP, RP]
        l.orc_null_space.argtypes = [C.c_void_p]
        l.orc_factor_step.argtypes = [C.c_void_p]
        l.orc_update_state.argtypes = [C.c_void_p, RP, RP, RP]
        l.orc_eliminate.argtypes = [C.c_void_p, RP, RP, C.c_int, C.c_int]
        for fn in ("orc_apg_init", "orc_solve_step", "orc_prox", "orc_residual", "orc_dual_update"):
            getattr(l, fn).argtypes = [C.c_void_p]
        l.orc_extrapolate.argtypes = [C.c_void_p, cf]
        l.orc_primal_infeasibility.argtypes = [C.c_void_p]
        l.orc_primal_infeasibility.restype = cf
        l.orc_lambda_table.argtypes = [C.c_int, RP]
        l.orc_apg.argtypes = [C.c_void_p, C.c_int, RP]
        l.orc_count.argtypes = [C.c_void_p, C.c_char_p]
        l.orc_count.restype = C.c_long
        l.orc_get.argtypes = [C.c_void_p, C.c_char_p, RP, C.c_long]
        l.orc_set.argtypes = [C.c_void_p, C.c_char_p, RP, C.c_long]
        l.orc_final_branch_node.argtypes = [C.c_void_p]
        l.orc_distance.argtypes = [C.c_void_p, C.c_int]
        l.orc_distance.restype = cf
        l.orc_set_num_threads.argtypes = [C.c_int]

    def arr(self, a):
        return np.ascontiguousarray(a, dtype=self.np)

    def ptr(self, a):
        return a.ctypes.data_as(self.RP)


_libs = {}


def lib(f64: bool = False) -> _Lib:
    if f64 not in _libs:
        build()
        _libs[f64] = _Lib(_LIB_PATH64 if f64 else _LIB_PATH, f64)
    return _libs[f64]


def _ip(a):
    return a.ctypes.data_as(IP)


def lambda_table(iters: int) -> np.ndarray:
    L = lib()
    out = np.zeros(iters, dtype=np.float32)
    L.l.orc_lambda_table(iters, L.ptr(out))
    return out


class Oracle:
    """One problem instance on the CPU oracle.  `problem` is a rapidnet_b200.problem.Problem.
    precision="f32" is the oracle (fp32 like the reference); "f64" runs the same code in double."""

    def __init__(self, problem, L=None, Lhat=None, threads: int | None = None, precision: str = "f32"):
        self.L_ = L_ = lib(precision == "f64")
        l = L_.l
        if threads:
            l.orc_set_num_threads(int(threads))
        n, t, c = problem.network, problem.tree, problem.config
        rl = L_.arr
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = [i32(t.stages), i32(t.nodes_per_stage), i32(t.nodes_per_stage_cumul), i32(t.ancestor),
                      i32(t.n_children), i32(t.n_children_cumul), rl(t.prob), rl(t.err_demand), rl(t.err_price),
                      rl(n.B), rl(n.Gd), rl(n.E), rl(n.Ed), rl(n.xmin), rl(n.xmax), rl(n.xsafe),
                      rl(n.umin), rl(n.umax), rl(n.alpha1), rl(c.costW), rl(c.precond)]
        k = self._keep
        fp = L_.ptr
        self.problem = problem
        self.h = l.orc_create(n.nx, n.nu, n.nd, n.ne, c.nv, t.N, t.K, t.nodes, t.n_nonleaf,
                              *[_ip(a) for a in k[:6]], *[fp(a) for a in k[6:9]],
                              *[fp(a) for a in k[9:19]], fp(k[19]), fp(k[20]),
                              float(np.float32(c.penalty_x)), float(np.float32(c.penalty_xs)), float(np.float32(c.step_size)))
        if L is not None:
            self.set_L(L, Lhat)

    def close(self):
        if self.h:
            self.L_.l.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_L(self, L, Lhat):
        L, Lhat = self.L_.arr(L), self.L_.arr(Lhat)
        self.L_.l.orc_set_L(self.h, self.L_.ptr(L), self.L_.ptr(Lhat))

    def factor_step(self):
        rc = self.L_.l.orc_factor_step(self.h)
        if rc:
            raise RuntimeError(f"oracle factor step failed rc={rc}")

    def update_state(self, x=None, uprev=None, dprev=None):
        c, a, p = self.problem.config, self.L_.arr, self.L_.ptr
        x = a(c.current_x if x is None else x)
        u = a(c.prev_u if uprev is None else uprev)
        d = a(c.prev_demand if dprev is None else dprev)
        self.L_.l.orc_update_state(self.h, p(x), p(u), p(d))

    def eliminate(self, dhat, alphahat, demand_uncertainty=True, price_uncertainty=True):
        dhat, ah = self.L_.arr(dhat), self.L_.arr(alphahat)
        self.L_.l.orc_eliminate(self.h, self.L_.ptr(dhat), self.L_.ptr(ah), int(demand_uncertainty), int(price_uncertainty))

    def apg_init(self): self.L_.l.orc_apg_init(self.h)
    def extrapolate(self, lam): self.L_.l.orc_extrapolate(self.h, float(lam))
    def solve_step(self): self.L_.l.orc_solve_step(self.h)
    def prox(self): self.L_.l.orc_prox(self.h)
    def residual(self): self.L_.l.orc_residual(self.h)
    def dual_update(self): self.L_.l.orc_dual_update(self.h)
    def primal_infeasibility(self): return float(self.L_.l.orc_primal_infeasibility(self.h))
    def distance(self, which): return float(self.L_.l.orc_distance(self.h, which))

    def apg(self, iters: int) -> np.ndarray:
        infs = np.zeros(iters, dtype=self.L_.np)
        self.L_.l.orc_apg(self.h, iters, self.L_.ptr(infs))
        return infs

    def get(self, name: str) -> np.ndarray:
        cnt = self.L_.l.orc_count(self.h, name.encode())
        if cnt < 0:
            raise KeyError(name)
        out = np.zeros(cnt, dtype=self.L_.np)
        rc = self.L_.l.orc_get(self.h, name.encode(), self.L_.ptr(out), cnt)
        assert rc == 0
        return out

    def set(self, name: str, arr):
        arr = self.L_.arr(arr).reshape(-1)
        rc = self.L_.l.orc_set(self.h, name.encode(), self.L_.ptr(arr), arr.size)
        if rc:
            raise KeyError(f"{name} (count {arr.size})")

    @property
    def final_branch_node(self):
        return self.L_.l.orc_final_branch_node(self.h)
