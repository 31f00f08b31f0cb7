"""Debug helper: phase clock of the persistent kernel (no oracle)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapidnet_b200 import cabi
from rapidnet_b200.datagen import named_problem
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
t0 = time.time()
prob = named_problem(name)
print("problem", time.time() - t0, flush=True)
s = cabi.Solver(prob)
s.set_modes(cabi.SWEEP_PERSISTENT, cabi.FACTORS_DF if "df" in sys.argv else cabi.FACTORS_FULL)
s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[0], prob.forecast.prices[0]); s.sync()
print("setup", time.time() - t0, flush=True)
for n in (1, 2, 20):
    s.apg_solve(n); print("solve", n, time.time() - t0, flush=True)
print(name, "profile", s.profile_kernels(100), flush=True)
print("phases /iter", {k: round(v) for k, v in s.phase_times().items()})
ct = s.cta_times()
if ct:
    import numpy as np
    t = np.array([c[0] for c in ct]); sm = np.array([c[1] for c in ct])
    order = np.argsort(t)
    print("phase S per CTA: min %.0f median %.0f max %.0f ns" % (t.min(), np.median(t), t.max()))
    print("slowest (cta, sm, ns):", [(int(k), int(sm[k]), int(t[k])) for k in order[-12:]])
    print("fastest (cta, sm, ns):", [(int(k), int(sm[k]), int(t[k])) for k in order[:12]])
