"""GPU: the shared-factor formulation (RN_FACTORS_SHARED, SURVEY.md 8f-4 "Tier B") against the CPU oracle.

No per-node Engine matrix is read: D_i xi = G (sysF_i' xi), F_i psi = L' (s_u o psi) (Engine.cu:720-728 builds D_i, F_i as
exactly those products) and v = -1/2 Omega r.  Same iterates as the reference formulation up to fp32 rounding, so the
same tolerance as tests/test_gpu_parity.py: relative 1e-4, widened to KAPPA x the fp32 noise floor where that floor
exceeds 1e-4 (refcompare.floor_tol)."""
import numpy as np
import pytest

from rapidnet_b200 import cabi
from rapidnet_b200.datagen import named_problem
from refcompare import RTOL, floor_tol, pinf_close, rel_err
from test_gpu_parity import _check_u0, _compare_state, _setup, _setup64

pytestmark = pytest.mark.gpu


def test_toy_shared_factors_match_oracle(toy):
    prob = toy[0]
    s, o = _setup(prob, cabi.SWEEP_PERSISTENT, cabi.FACTORS_SHARED, slot=1)
    o64 = _setup64(prob, s, slot=1)
    for iters in (1, 10, 100, 500):
        u0, infs = s.apg_solve(iters, want_infs=True)
        oinfs = o.apg(iters)
        o64.apg(iters)
        _compare_state(s, o, f"toy shared it={iters}", o64)
        _check_u0(u0, o, o64, f"toy shared it={iters}")
        pinf_close(infs, oinfs)
    s.close(); o.close(); o64.close()


@pytest.mark.parametrize("name", ["C1", "C1r6", "C1r30"])
def test_barcelona_shared_factors_match_oracle(name):
    prob = named_problem(name)
    s, o = _setup(prob, cabi.SWEEP_PERSISTENT, cabi.FACTORS_SHARED)
    assert s.info().factor_mode == cabi.FACTORS_SHARED
    o64 = _setup64(prob, s)
    for iters in (1, 10, 100, 500):
        u0, _ = s.apg_solve(iters)
        o.apg(iters)
        o64.apg(iters)
        worst = _compare_state(s, o, f"{name} shared it={iters}", o64)
        _check_u0(u0, o, o64, f"{name} shared it={iters}")
        print(f"{name} shared it={iters}: worst rel err {worst:.2e}")
    s.close(); o.close(); o64.close()


def test_shared_factors_agree_with_streamed_on_c2():
    """BASELINE config[1] at full size: the shared-factor solve against the streamed (Engine-matrix) solve of the same
    handle after 60 iterations, and the size-independent dynamics identity on its own output."""
    prob = named_problem("C2", max_iter=60)
    s = cabi.Solver(prob)
    s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[0], prob.forecast.prices[0])
    s.apg_solve(60)
    ref = {k: s.read(k) for k in ("VEC_U", "VEC_X", "VEC_UPDATE_XI", "VEC_UPDATE_PSI", "VEC_DUAL_XI")}
    s.set_modes(cabi.SWEEP_PERSISTENT, cabi.FACTORS_SHARED)
    s.apg_solve(60)
    for k, want in ref.items():
        assert rel_err(s.read(k), want) < RTOL, k
    n, t = prob.network, prob.tree
    X = s.read("VEC_X").reshape(-1, n.nx); U = s.read("VEC_U").reshape(-1, n.nu); e = s.read("VEC_E").reshape(-1, n.nx)
    B = n.B.reshape(n.nu, n.nx).T
    par = t.ancestor.astype(np.int64) - 1
    xp = np.where(par[:, None] >= 0, X[np.maximum(par, 0)], prob.config.current_x[None, :])
    assert rel_err(X, xp + e + U @ B.T) < 1e-5
    info = s.info()
    # Tier-B bytes: no per-node matrix term in what one iteration has to move
    assert info.stream_bytes_per_iteration < 0.1 * info.factor_bytes
    s.close()


def test_shared_factors_need_the_persistent_sweep(toy):
    prob = toy[0]
    s = cabi.Solver(prob)
    s.set_modes(cabi.SWEEP_CHAIN, cabi.FACTORS_SHARED)
    s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[1], prob.forecast.prices[1])
    with pytest.raises(RuntimeError):
        s.apg_solve(5)
    s.close()
