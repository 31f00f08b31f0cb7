"""One cold-started solve of a named workload through the C ABI (for ncu captures): python tools/ncu_solve.py C2 10 [factors]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapidnet_b200 import cabi  # noqa: E402
from rapidnet_b200.datagen import named_problem  # noqa: E402

name, iters = sys.argv[1], int(sys.argv[2])
factors = {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[sys.argv[3] if len(sys.argv) > 3 else "full"]
prob = named_problem(name, max_iter=iters)
s = cabi.Solver(prob)
s.set_modes(cabi.SWEEP_PERSISTENT, factors)
s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[0], prob.forecast.prices[0])
s.apg_solve(iters)      # warm-up launch (allocations, L2 state)
s.apg_solve(iters)      # the launch to capture (ncu -k regex:k_apg_persistent --launch-skip 1 -c 1)
s.sync()
print("done", name, iters, s.info().sweep_mode)
s.close()
