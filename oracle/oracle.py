"""ctypes wrapper of oracle/rapidnet_oracle.c (test infrastructure, see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librapidnet_oracle.so")
_lib = None

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rapidnet_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int] * 9 + [IP] * 6 + [FP] * 3 + [FP] * 10 + [FP, FP, C.c_float, C.c_float, C.c_float]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_set_L.argtypes = [C.c_void_p, FP, FP]
        _lib.orc_null_space.argtypes = [C.c_void_p]
        _lib.orc_factor_step.argtypes = [C.c_void_p]
        _lib.orc_update_state.argtypes = [C.c_void_p, FP, FP, FP]
        _lib.orc_eliminate.argtypes = [C.c_void_p, FP, FP, C.c_int, C.c_int]
        for fn in ("orc_apg_init", "orc_solve_step", "orc_prox", "orc_residual", "orc_dual_update"):
            getattr(_lib, fn).argtypes = [C.c_void_p]
        _lib.orc_extrapolate.argtypes = [C.c_void_p, C.c_float]
        _lib.orc_primal_infeasibility.argtypes = [C.c_void_p]
        _lib.orc_primal_infeasibility.restype = C.c_float
        _lib.orc_lambda_table.argtypes = [C.c_int, FP]
        _lib.orc_apg.argtypes = [C.c_void_p, C.c_int, FP]
        _lib.orc_count.argtypes = [C.c_void_p, C.c_char_p]
        _lib.orc_count.restype = C.c_long
        _lib.orc_get.argtypes = [C.c_void_p, C.c_char_p, FP, C.c_long]
        _lib.orc_set.argtypes = [C.c_void_p, C.c_char_p, FP, C.c_long]
        _lib.orc_final_branch_node.argtypes = [C.c_void_p]
        _lib.orc_distance.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_distance.restype = C.c_float
        _lib.orc_set_num_threads.argtypes = [C.c_int]
    return _lib


def _fp(a):
    return a.ctypes.data_as(FP)


def _ip(a):
    return a.ctypes.data_as(IP)


def lambda_table(iters: int) -> np.ndarray:
    out = np.zeros(iters, dtype=np.float32)
    lib().orc_lambda_table(iters, _fp(out))
    return out


class Oracle:
    """One problem instance on the CPU oracle.  `problem` is a rapidnet_b200.problem.Problem."""

    def __init__(self, problem, L=None, Lhat=None, threads: int | None = None):
        l = lib()
        if threads:
            l.orc_set_num_threads(int(threads))
        n, t, c = problem.network, problem.tree, problem.config
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = [i32(t.stages), i32(t.nodes_per_stage), i32(t.nodes_per_stage_cumul), i32(t.ancestor),
                      i32(t.n_children), i32(t.n_children_cumul), f32(t.prob), f32(t.err_demand), f32(t.err_price),
                      f32(n.B), f32(n.Gd), f32(n.E), f32(n.Ed), f32(n.xmin), f32(n.xmax), f32(n.xsafe),
                      f32(n.umin), f32(n.umax), f32(n.alpha1), f32(c.costW), f32(c.precond)]
        k = self._keep
        self.problem = problem
        self.h = l.orc_create(n.nx, n.nu, n.nd, n.ne, c.nv, t.N, t.K, t.nodes, t.n_nonleaf,
                              *[_ip(a) for a in k[:6]], *[_fp(a) for a in k[6:9]],
                              *[_fp(a) for a in k[9:19]], _fp(k[19]), _fp(k[20]),
                              float(c.penalty_x), float(c.penalty_xs), float(c.step_size))
        if L is not None:
            self.set_L(L, Lhat)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_L(self, L, Lhat):
        L = np.ascontiguousarray(L, dtype=np.float32)
        Lhat = np.ascontiguousarray(Lhat, dtype=np.float32)
        lib().orc_set_L(self.h, _fp(L), _fp(Lhat))

    def factor_step(self):
        rc = lib().orc_factor_step(self.h)
        if rc:
            raise RuntimeError(f"oracle factor step failed rc={rc}")

    def update_state(self, x=None, uprev=None, dprev=None):
        c = self.problem.config
        x = np.ascontiguousarray(c.current_x if x is None else x, dtype=np.float32)
        u = np.ascontiguousarray(c.prev_u if uprev is None else uprev, dtype=np.float32)
        d = np.ascontiguousarray(c.prev_demand if dprev is None else dprev, dtype=np.float32)
        lib().orc_update_state(self.h, _fp(x), _fp(u), _fp(d))

    def eliminate(self, dhat, alphahat, demand_uncertainty=True, price_uncertainty=True):
        dhat = np.ascontiguousarray(dhat, dtype=np.float32)
        ah = np.ascontiguousarray(alphahat, dtype=np.float32)
        lib().orc_eliminate(self.h, _fp(dhat), _fp(ah), int(demand_uncertainty), int(price_uncertainty))

    def apg_init(self): lib().orc_apg_init(self.h)
    def extrapolate(self, lam): lib().orc_extrapolate(self.h, float(lam))
    def solve_step(self): lib().orc_solve_step(self.h)
    def prox(self): lib().orc_prox(self.h)
    def residual(self): lib().orc_residual(self.h)
    def dual_update(self): lib().orc_dual_update(self.h)
    def primal_infeasibility(self): return float(lib().orc_primal_infeasibility(self.h))
    def distance(self, which): return float(lib().orc_distance(self.h, which))

    def apg(self, iters: int) -> np.ndarray:
        infs = np.zeros(iters, dtype=np.float32)
        lib().orc_apg(self.h, iters, _fp(infs))
        return infs

    def get(self, name: str) -> np.ndarray:
        cnt = lib().orc_count(self.h, name.encode())
        if cnt < 0:
            raise KeyError(name)
        out = np.zeros(cnt, dtype=np.float32)
        rc = lib().orc_get(self.h, name.encode(), _fp(out), cnt)
        assert rc == 0
        return out

    def set(self, name: str, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        rc = lib().orc_set(self.h, name.encode(), _fp(arr), arr.size)
        if rc:
            raise KeyError(f"{name} (count {arr.size})")

    @property
    def final_branch_node(self):
        return lib().orc_final_branch_node(self.h)
