"""GPU: the reference's own test suite, replayed through the C ABI against its golden vectors.

Mirrors testEngineTesting (/root/reference/src/test/Testing.cu:340-477) and testSmpcController
(:482-531 -> TestSmpcController.cu:114-398) with the reference's tolerances (tests/refcompare.py)."""
import numpy as np
import pytest

from rapidnet_b200 import cabi
from refcompare import engine_close, smpc_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[cabi.SWEEP_PERSISTENT, cabi.SWEEP_CHAIN, cabi.SWEEP_PER_STAGE, cabi.SWEEP_BATCHED],
                ids=["persistent", "chain", "per_stage", "batched"])
def solver(request, toy):
    prob, engine, _ = toy
    s = cabi.Solver(prob)
    s.set_modes(request.param, cabi.FACTORS_FULL)
    # golden factor matrices are in MATLAB's null-space basis (engineTest.json: matL); SURVEY 7.3-5
    s.set_null_space(engine["matL"], prob.config.Lhat)
    s.factor_step()
    s.update_state()
    s.eliminate_coupling(prob.forecast.demand[1], prob.forecast.prices[1])    # timeInst = 1
    yield s
    s.close()


def _nodes(s, name, dim, nodes_1based, count):
    full = s.read(name).reshape(-1, dim)
    return np.concatenate([full[int(n) - 1] for n in nodes_1based[:count]])


def test_engine_golden(solver, toy):
    prob, g, _ = toy
    s = solver
    nx, nu, nv, N = prob.network.nx, prob.network.nu, prob.config.nv, prob.tree.N
    engine_close(s.read("VEC_UHAT"), g["uHat"])
    engine_close(s.read("VEC_E"), g["vecE"])
    engine_close(s.read("VEC_BETA"), g["beta"])
    engine_close(s.read("SYS_MAT_L"), g["matL"])
    engine_close(s.read("VEC_ALPHA"), g["costAlpha"])
    sn = g["scenarioNodes"]
    fbs = s.info().final_branch_stage
    assert fbs == prob.tree.final_branch_stage() and s.info().final_branch_node == prob.tree.final_branch_node()
    for name, key, dim in (("SYS_MAT_F", "sysF", 2 * nx * nx), ("SYS_MAT_G", "sysG", nu * nu),
                           ("SYS_XMIN", "xmin", nx), ("SYS_XMAX", "xmax", nx), ("SYS_XS", "xs", nx),
                           ("SYS_UMIN", "umin", nu), ("SYS_UMAX", "umax", nu),
                           ("MAT_D", "d", 2 * nx * nv), ("MAT_F", "f", nu * nv),
                           ("MAT_PHI", "Phi", 2 * nx * nv), ("MAT_PSI", "Psi", nu * nv)):
        engine_close(_nodes(s, name, dim, sn, N), g[key])
    engine_close(_nodes(s, "MAT_OMEGA", nv * nv, sn, fbs), g["omega"][: fbs * nv * nv])
    engine_close(_nodes(s, "MAT_THETA", nx * nv, sn, fbs), g["Theta"][: fbs * nx * nv])
    engine_close(np.concatenate([s.read("MAT_G")] * fbs), g["g"][: fbs * nx * nv])
    xsu = s.read("SYS_XS_UPPER")
    assert np.all(xsu.view(np.uint32) == 0x7F7F7F7F)


def test_apg_steps_golden(solver, toy):
    _, _, g = toy
    s = solver
    s.apg_init()
    # testExtrapolation
    s.write("VEC_XI", g["xi"]); s.write("VEC_PSI", g["psi"])
    s.write("VEC_UPDATE_XI", g["updateXi"]); s.write("VEC_UPDATE_PSI", g["updatePsi"])
    th = g["theta"].astype(np.float32)
    lam = np.float32(th[1] * (np.float32(1) / th[0] - np.float32(1)))
    s.step(cabi.STEP_EXTRAPOLATE, lam)
    smpc_close(s.read("VEC_ACCEL_XI"), g["acceleXi"]); smpc_close(s.read("VEC_ACCEL_PSI"), g["accelePsi"])
    smpc_close(s.read("VEC_XI"), g["finalXi"]); smpc_close(s.read("VEC_PSI"), g["finalPsi"])
    # testSoveStep
    s.write("VEC_ACCEL_XI", g["acceleXi"]); s.write("VEC_ACCEL_PSI", g["accelePsi"])
    s.step(cabi.STEP_SOLVE)
    smpc_close(s.read("VEC_X"), g["X"]); smpc_close(s.read("VEC_U"), g["U"])
    engine_close(s.read("VEC_V"), g["tempV"], tol=1e-3)
    # testProximalStep
    s.write("VEC_X", g["X"]); s.write("VEC_U", g["U"])
    s.step(cabi.STEP_PROX)
    smpc_close(s.read("VEC_PRIMAL_XI"), g["primalX"]); smpc_close(s.read("VEC_PRIMAL_PSI"), g["primalU"])
    smpc_close(s.read("VEC_DUAL_XI"), g["dualX"]); smpc_close(s.read("VEC_DUAL_PSI"), g["dualU"])
    # testFixedPointResidual
    s.write("VEC_PRIMAL_XI", g["primalX"]); s.write("VEC_PRIMAL_PSI", g["primalU"])
    s.write("VEC_DUAL_XI", g["dualX"]); s.write("VEC_DUAL_PSI", g["dualU"])
    s.step(cabi.STEP_RESIDUAL)
    smpc_close(s.read("VEC_RESIDUAL_XI"), g["fixedPointResidualXi"])
    smpc_close(s.read("VEC_RESIDUAL_PSI"), g["fixedPointResidualPsi"])
    # testDualUpdate
    s.write("VEC_RESIDUAL_XI", g["fixedPointResidualXi"]); s.write("VEC_RESIDUAL_PSI", g["fixedPointResidualPsi"])
    s.step(cabi.STEP_DUAL_UPDATE)
    smpc_close(s.read("VEC_UPDATE_XI"), g["finalUpdateXi"]); smpc_close(s.read("VEC_UPDATE_PSI"), g["finalUpdatePsi"])


def test_persistent_kernel_golden_iteration(toy):
    """The golden step vectors through k_apg_persistent itself (rn_step dispatches to the stand-alone kernels): ONE fused
    iteration continued from the golden y_k / y_{k-1} with the golden lambda must reproduce the extrapolation, the solve
    step, the prox, the residual and the dual update of TestSmpcController.cu:134-398 at the reference's tolerances."""
    prob, engine, g = toy
    s = cabi.Solver(prob)
    s.set_modes(cabi.SWEEP_PERSISTENT, cabi.FACTORS_FULL)
    s.set_null_space(engine["matL"], prob.config.Lhat)
    s.factor_step(); s.update_state()
    s.eliminate_coupling(prob.forecast.demand[1], prob.forecast.prices[1])
    assert s.info().sweep_mode == cabi.SWEEP_PERSISTENT
    s.apg_init()
    s.write("VEC_XI", g["xi"]); s.write("VEC_PSI", g["psi"])
    s.write("VEC_UPDATE_XI", g["updateXi"]); s.write("VEC_UPDATE_PSI", g["updatePsi"])
    th = g["theta"].astype(np.float32)
    lam = np.float32(th[1] * (np.float32(1) / th[0] - np.float32(1)))
    launches = s.info().kernel_launches
    s.apg_continue(1, lambdas=[lam])
    assert s.info().kernel_launches - launches == 5      # 3 chain-major copies + k_apg_persistent + k_finalize
    smpc_close(s.read("VEC_ACCEL_XI"), g["acceleXi"]); smpc_close(s.read("VEC_ACCEL_PSI"), g["accelePsi"])
    smpc_close(s.read("VEC_XI"), g["finalXi"]); smpc_close(s.read("VEC_PSI"), g["finalPsi"])     # y_{k-1} <- y_k
    smpc_close(s.read("VEC_X"), g["X"]); smpc_close(s.read("VEC_U"), g["U"])
    engine_close(s.read("VEC_V"), g["tempV"], tol=1e-3)
    smpc_close(s.read("VEC_PRIMAL_XI"), g["primalX"]); smpc_close(s.read("VEC_PRIMAL_PSI"), g["primalU"])
    smpc_close(s.read("VEC_DUAL_XI"), g["dualX"]); smpc_close(s.read("VEC_DUAL_PSI"), g["dualU"])
    smpc_close(s.read("VEC_RESIDUAL_XI"), g["fixedPointResidualXi"])
    smpc_close(s.read("VEC_RESIDUAL_PSI"), g["fixedPointResidualPsi"])
    # the golden dual update starts from the golden residual and the golden w (TestSmpcController.cu:369-389), both
    # reproduced above, so y_{k+1} = w + step * res must match finalUpdate as well
    smpc_close(s.read("VEC_UPDATE_XI"), g["finalUpdateXi"]); smpc_close(s.read("VEC_UPDATE_PSI"), g["finalUpdatePsi"])
    s.close()


def test_warm_start_continues_the_iteration(toy):
    """rn_apg_continue with the theta sequence of a longer solve continues it exactly: 30 cold iterations followed by 20
    continued ones (with lambda_30..49) equal 50 cold iterations; warm_restart restarts theta from the final duals."""
    prob = toy[0]
    s = cabi.Solver(prob)
    s.factor_step(); s.update_state(); s.eliminate_coupling(prob.forecast.demand[1], prob.forecast.prices[1])
    s.apg_solve(50)
    want = {k: s.read(k) for k in ("VEC_U", "VEC_X", "VEC_UPDATE_XI", "VEC_UPDATE_PSI", "VEC_XI")}
    th0, th1, lams = np.float32(1), np.float32(1), []
    for _ in range(50):
        lams.append(np.float32(th1 * (np.float32(1) / th0 - np.float32(1))))
        th0 = th1
        th1 = np.float32(0.5 * (np.sqrt(float(th1) ** 4 + 4 * float(th1) ** 2) - float(th1) ** 2))
    s.apg_solve(30)
    s.apg_continue(20, lambdas=lams[30:])
    for k, v in want.items():
        assert np.allclose(s.read(k), v, rtol=2e-5, atol=1e-3 * max(1.0, float(np.abs(v).max()))), k
    y = s.read("VEC_UPDATE_XI").copy()
    s.apg_continue(1, warm_restart=True)
    assert np.array_equal(s.read("VEC_XI"), y)           # y_{-1} = y_0 = the previous duals
    assert np.isfinite(s.read("VEC_U")).all()
    s.close()


def test_cusolver_null_space(toy):
    """Default path (Engine::calculateMatLandMatLhat through cuSOLVER Dgesvd): a valid orthonormal basis of
    null(E) spanning the golden subspace, and the basis-free Lhat of the shipped config."""
    prob, g, _ = toy
    s = cabi.Solver(prob)
    s.factor_step()
    nu, nv, nd, ne = prob.network.nu, prob.config.nv, prob.network.nd, prob.network.ne
    L = s.read("SYS_MAT_L").reshape(nv, nu).T
    Lg = np.asarray(g["matL"]).reshape(nv, nu).T
    E = prob.network.E.reshape(nu, ne).T
    assert np.abs(E @ L).max() < 1e-5
    assert np.abs(L.T @ L - np.eye(nv)).max() < 1e-5
    assert np.abs(L @ L.T - Lg @ Lg.T).max() < 1e-5
    Lhat = s.read("SYS_MAT_LHAT").reshape(nd, nu).T
    assert np.abs(Lhat - prob.config.Lhat.reshape(nd, nu).T).max() < 1e-5
    # basis-invariant outputs still meet the golden vectors
    s.update_state(); s.eliminate_coupling(prob.forecast.demand[1], prob.forecast.prices[1])
    engine_close(s.read("VEC_UHAT"), g["uHat"]); engine_close(s.read("VEC_E"), g["vecE"])
    s.close()
