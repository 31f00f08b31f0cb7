"""Debug helper: persistent kernel vs the CPU oracle, per-buffer relative errors (no asserts)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import Oracle
from rapidnet_b200 import cabi
from rapidnet_b200.datagen import named_problem
from rapidnet_b200.problem import problem_from_npz_dict

PAIRS = [("VEC_U", "U"), ("VEC_X", "X"), ("VEC_V", "V"), ("VEC_UPDATE_XI", "update_xi"), ("VEC_UPDATE_PSI", "update_psi"),
         ("VEC_XI", "xi"), ("VEC_PSI", "psi"), ("VEC_DUAL_XI", "dual_xi"), ("VEC_DUAL_PSI", "dual_psi"),
         ("VEC_PRIMAL_XI", "primal_xi"), ("VEC_PRIMAL_PSI", "primal_psi"), ("VEC_ACCEL_XI", "accel_xi"),
         ("VEC_ACCEL_PSI", "accel_psi"), ("VEC_RESIDUAL_XI", "res_xi"), ("VEC_RESIDUAL_PSI", "res_psi")]

def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))

name = sys.argv[1] if len(sys.argv) > 1 else "toy"
factors = cabi.FACTORS_DF if "df" in sys.argv else cabi.FACTORS_FULL
if name == "toy":
    z = dict(np.load("tests/golden/toy.npz", allow_pickle=False)); prob = problem_from_npz_dict(z); slot = 1
else:
    prob = named_problem(name); slot = 0
sweep = cabi.SWEEP_CHAIN if "chain" in sys.argv else (cabi.SWEEP_PER_STAGE if "per_stage" in sys.argv else cabi.SWEEP_PERSISTENT)
s = cabi.Solver(prob); s.set_modes(sweep, factors)
s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[slot], prob.forecast.prices[slot])
o = Oracle(prob, L=s.read("SYS_MAT_L"), Lhat=s.read("SYS_MAT_LHAT"))
o.factor_step(); o.update_state(); o.eliminate(prob.forecast.demand[slot], prob.forecast.prices[slot])
print("info: chain stage", s.info().chain_first_stage, "sweep", s.info().sweep_mode, flush=True)
for iters in [int(x) for x in os.environ.get("DBG_ITERS", "1,10,100,500").split(",")]:
    u0, infs = s.apg_solve(iters, want_infs=True)
    oinf = o.apg(iters)
    errs = {g: rel(s.read(g), o.get(n)) for g, n in PAIRS}
    bad = {k: f"{v:.1e}" for k, v in errs.items() if not v < 1e-4}
    print(f"{name} it={iters}: worst {max(errs.values()):.2e} bad={bad} pinf max diff {np.abs(infs-oinf).max():.3e} at {int(np.abs(infs-oinf).argmax())} ({infs[int(np.abs(infs-oinf).argmax())]:.4e} vs {oinf[int(np.abs(infs-oinf).argmax())]:.4e})", flush=True)
print("profile", s.profile_kernels(50))
print("phases ns/iter", {k: round(v) for k, v in s.phase_times().items()})
