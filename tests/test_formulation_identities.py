"""The algebra behind RN_FACTORS_DF and RN_FACTORS_SHARED, checked on the CPU against the reference's OWN golden vectors
(src/test/testDataFiles/engineTest.json -> tests/golden/toy.npz, 24 sampled nodes) and against the oracle at Barcelona
size:
    sysF_i = [diag(s_x); diag(s_xs)],  sysG_i = diag(s_u)                      (Engine.cu:382-463)
    D_i = G sysF_i',  F_i = L' sysG_i'                                          (Engine.cu:720-728)
    Phi_i = -1/2 Omega_i D_i,  Psi_i = -1/2 Omega_i F_i,  Theta_i = -1/2 Omega_i G  (Engine.cu:729-747)
so that  D_i xi = G (sysF_i' xi),  F_i psi = L' (s_u o psi)  and  v_i = -1/2 Omega_i r_i  -- no per-node matrix is needed.
All matrices are column-major as the reference stores them."""
import numpy as np

from oracle.oracle import Oracle
from rapidnet_b200.datagen import named_problem


def _cm(a, rows, cols):
    return np.asarray(a, dtype=np.float64).reshape(cols, rows).T


def test_identities_hold_in_the_reference_golden_vectors(toy):
    prob, eng, _ = toy
    nx, nu, nv = prob.network.nx, prob.network.nu, prob.dims["nv"]
    n = eng["scenarioNodes"].size
    L = _cm(eng["matL"], nu, nv)
    for i in range(n):
        D = _cm(eng["d"][i * nv * 2 * nx:(i + 1) * nv * 2 * nx], nv, 2 * nx)
        F = _cm(eng["f"][i * nv * nu:(i + 1) * nv * nu], nv, nu)
        G = _cm(eng["g"][i * nv * nx:(i + 1) * nv * nx], nv, nx)
        Om = _cm(eng["omega"][i * nv * nv:(i + 1) * nv * nv], nv, nv)
        Phi = _cm(eng["Phi"][i * nv * 2 * nx:(i + 1) * nv * 2 * nx], nv, 2 * nx)
        Psi = _cm(eng["Psi"][i * nv * nu:(i + 1) * nv * nu], nv, nu)
        Th = _cm(eng["Theta"][i * nv * nx:(i + 1) * nv * nx], nv, nx)
        sysF = _cm(eng["sysF"][i * 2 * nx * nx:(i + 1) * 2 * nx * nx], 2 * nx, nx)
        sysG = _cm(eng["sysG"][i * nu * nu:(i + 1) * nu * nu], nu, nu)
        # the system matrices are stacked diagonals
        assert np.count_nonzero(sysF[:nx] - np.diag(np.diag(sysF[:nx]))) == 0
        assert np.count_nonzero(sysF[nx:] - np.diag(np.diag(sysF[nx:]))) == 0
        assert np.count_nonzero(sysG - np.diag(np.diag(sysG))) == 0
        tol = dict(rtol=2e-3, atol=1e-2)        # the fixtures are printed with a few digits (Testing.cu compares at 1e-2)
        assert np.allclose(D, G @ sysF.T, **tol), i
        assert np.allclose(F, L.T @ sysG.T, **tol), i
        assert np.allclose(Phi, -0.5 * Om @ D, **tol), i
        assert np.allclose(Psi, -0.5 * Om @ F, **tol), i
        assert np.allclose(Th, -0.5 * Om @ G, **tol), i
        # hence, for any dual block: D xi = G (sysF' xi) and F psi = L' (s_u o psi)
        rng = np.random.default_rng(i)
        xi, psi = rng.standard_normal(2 * nx), rng.standard_normal(nu)
        c = np.diag(sysF[:nx]) * xi[:nx] + np.diag(sysF[nx:]) * xi[nx:]
        assert np.allclose(D @ xi, G @ c, rtol=2e-3, atol=1e-2 * np.abs(D).max())
        assert np.allclose(F @ psi, L.T @ (np.diag(sysG) * psi), rtol=2e-3, atol=1e-2 * np.abs(F).max())


def test_identities_hold_in_the_oracle_at_barcelona_size():
    prob = named_problem("C1", max_iter=5)
    o = Oracle(prob, L=prob.config.L, Lhat=prob.config.Lhat, precision="f64")
    o.factor_step()
    nx, nu, nv, nodes = prob.network.nx, prob.network.nu, prob.dims["nv"], prob.dims["nodes"]
    G, L = _cm(o.get("G"), nv, nx), _cm(o.get("L"), nu, nv)
    sx, sxs, su = o.get("s_x").reshape(nodes, nx), o.get("s_xs").reshape(nodes, nx), o.get("s_u").reshape(nodes, nu)
    D, F = o.get("D").reshape(nodes, -1), o.get("F").reshape(nodes, -1)
    Phi, Psi = o.get("Phi").reshape(nodes, -1), o.get("Psi").reshape(nodes, -1)
    Om = o.get("Omega").reshape(-1, nv * nv)
    fb = o.final_branch_node
    cum = np.concatenate([[0], np.cumsum(prob.tree.nodes_per_stage)])
    for i in (0, 1, nodes // 2, nodes - 1):
        Di, Fi = _cm(D[i], nv, 2 * nx), _cm(F[i], nv, nu)
        assert np.allclose(Di, np.hstack([G * sx[i], G * sxs[i]]), rtol=1e-12, atol=1e-12)
        assert np.allclose(Fi, L.T * su[i], rtol=1e-12, atol=1e-12)
        s = int(prob.tree.stages[i]); j = i - cum[s]
        k = fb - prob.dims["K"] + j if fb <= cum[s] else i                    # Omega aliasing, Engine.cu:210-221
        Oi = _cm(Om[k], nv, nv)
        assert np.allclose(_cm(Phi[i], nv, 2 * nx), -0.5 * Oi @ Di, rtol=1e-9, atol=1e-12)
        assert np.allclose(_cm(Psi[i], nv, nu), -0.5 * Oi @ Fi, rtol=1e-9, atol=1e-12)
    o.close()
