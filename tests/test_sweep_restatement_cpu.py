"""CPU: the algebra of the persistent kernel's backward sweep (phases B and C), restated in numpy and
checked against the oracle's solveStep (the reference's recursion, /root/reference/src/SmpcController.cu:593-673) in double.

The reference runs, stage by stage from the leaves,  q = c + sum_children q_c,  sigma = beta + sum_children r_c,
r = ((sigma + D xi) + F psi) + G q_bar,  v = Theta q_bar - 1/2 Omega sigma + Phi xi + Psi psi   (q_bar = sum_children q_c).
The kernel never forms q or c = sysF' xi:  D_i = G sysF_i'  (Engine.cu:720-728) gives  G c_i = D_i xi_i =: a_i,  so
  * along a chain,  G q  is a running sum of the streamed products a_i  (chain_rscan),
  * for a crown node i (stage s_i above the chain stage cs), with g_h = G q_h and r_h of the chain heads h below it,
        G q_bar_i = sum_{crown j below i} a_j + sum_h g_h
        G QS_i    = sum_{crown j below i} (s_j - s_i - 1) a_j + (cs - 1 - s_i) sum_h g_h
        sigma_i   = beta_i + [ sum_h r_h + sum_{crown j below i} (beta_j + a_j + F psi_j) ] + G QS_i        (crown_sums)
  * v_i = -1/2 Omega_i (sigma_i + G q_bar_i) + Phi_i xi_i + Psi_i psi_i     (Theta_i = -1/2 Omega_i G: ThetaBar is never loaded).
"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from rapidnet_b200.datagen import named_problem
from rapidnet_b200.partition import chain_stage


def _cm(a, rows, cols):
    return np.asarray(a, dtype=np.float64).reshape(cols, rows).T


@pytest.mark.parametrize("case", ["C1r6", "C1r30"])
def test_subtree_sum_form_of_the_backward_sweep_equals_the_reference_recursion(case):
    prob = named_problem(case, max_iter=5)
    t, net = prob.tree, prob.network
    nx, nu, nv, nodes, K = net.nx, net.nu, prob.dims["nv"], prob.dims["nodes"], prob.dims["K"]
    o = Oracle(prob, L=prob.config.L, Lhat=prob.config.Lhat, precision="f64")
    o.factor_step(); o.update_state(); o.eliminate(prob.forecast.demand[0], prob.forecast.prices[0])
    o.apg(4)                       # some non-trivial duals
    o.extrapolate(0.5)             # w (accel) of the next iteration
    o.solve_step()                 # the reference's recursion -> V
    V_ref = o.get("V").reshape(nodes, nv)
    wxi, wpsi = o.get("accel_xi").reshape(nodes, 2 * nx), o.get("accel_psi").reshape(nodes, nu)
    beta = o.get("beta").reshape(nodes, nv)
    D, F = o.get("D").reshape(nodes, -1), o.get("F").reshape(nodes, -1)
    Phi, Psi = o.get("Phi").reshape(nodes, -1), o.get("Psi").reshape(nodes, -1)
    Om = o.get("Omega").reshape(-1, nv * nv)
    fb = o.final_branch_node
    cum = np.concatenate([[0], np.cumsum(t.nodes_per_stage)]).astype(int)
    stages = np.asarray(t.stages, dtype=int)
    par = np.asarray(t.ancestor, dtype=int) - 1
    a = np.stack([_cm(D[i], nv, 2 * nx) @ wxi[i] for i in range(nodes)])          # the streamed products: D xi_w = G c
    f = np.stack([_cm(F[i], nv, nu) @ wpsi[i] for i in range(nodes)])
    phi = np.stack([_cm(Phi[i], nv, 2 * nx) @ wxi[i] for i in range(nodes)])
    psi = np.stack([_cm(Psi[i], nv, nu) @ wpsi[i] for i in range(nodes)])

    def omega(i):
        s = stages[i]; j = i - cum[s]
        k = fb - K + j if fb <= cum[s] else i                                      # Omega aliasing, Engine.cu:210-221
        return _cm(Om[k], nv, nv)

    cs = chain_stage(t)
    assert 0 < cs < t.N
    T = t.N - cs
    # ---- chains: running sums from the leaf up (chain_rscan)
    gq_head, r_head = np.zeros((K, nv)), np.zeros((K, nv))
    V = np.zeros((nodes, nv))
    for j in range(K):
        grun, rrun = np.zeros(nv), np.zeros(nv)
        for s in range(t.N - 1, cs - 1, -1):
            i = cum[s] + j
            sg = beta[i] + rrun
            V[i] = -0.5 * omega(i) @ (sg + grun) + phi[i] + psi[i]
            rrun = ((sg + a[i]) + f[i]) + grun
            grun = a[i] + grun
        gq_head[j], r_head[j] = grun, rrun
    # ---- crown: every node from its subtree sums, independently (crown_sums)
    head_anc = np.zeros((K, cs), dtype=int)                                        # crown ancestors of each chain head, by stage
    for j in range(K):
        p = par[cum[cs] + j]
        while p >= 0:
            head_anc[j, stages[p]] = p
            p = par[p]
    crown_anc = {}                                                                 # crown ancestors of each crown node
    for i in range(cum[cs]):
        p, lst = par[i], []
        while p >= 0:
            lst.append(p); p = par[p]
        crown_anc[i] = lst
    for i in range(cum[cs]):
        si = stages[i]
        below = [k for k in range(cum[cs]) if i in crown_anc[k]]
        heads = [j for j in range(K) if head_anc[j, si] == i]
        g_sum = sum((gq_head[j] for j in heads), np.zeros(nv))
        r_sum = sum((r_head[j] for j in heads), np.zeros(nv))
        gqbar = sum((a[k] for k in below), np.zeros(nv)) + g_sum
        gqs = sum(((stages[k] - si - 1) * a[k] for k in below), np.zeros(nv)) + (cs - 1 - si) * g_sum
        sigma = beta[i] + (r_sum + sum((beta[k] + a[k] + f[k] for k in below), np.zeros(nv))) + gqs
        V[i] = -0.5 * omega(i) @ (sigma + gqbar) + phi[i] + psi[i]
    scale = np.abs(V_ref).max()
    assert np.abs(V - V_ref).max() <= 1e-9 * scale, (case, float(np.abs(V - V_ref).max()), float(scale), T)
    o.close()


@pytest.mark.parametrize("case", ["C1r6", "C1r30"])
def test_path_sum_form_of_the_forward_sweep_equals_the_reference_recursion(case):
    """Phase F: the reference walks the stages,  u_i = uhat_i + (u_par - uhat_par) + L v_i,  x_i = x_par + e_i + B u_i
    (SmpcController.cu:675-741); the kernel gives every node  x_i = x_cur + sum_path e + B sum_path u  -- one column of a
    product with B per node, so crown and chains run in one phase (crown_forward, chain_uscan / chain_xscan)."""
    prob = named_problem(case, max_iter=5)
    t, net, c = prob.tree, prob.network, prob.config
    nx, nu, nv, nodes = net.nx, net.nu, prob.dims["nv"], prob.dims["nodes"]
    o = Oracle(prob, L=c.L, Lhat=c.Lhat, precision="f64")
    o.factor_step(); o.update_state(); o.eliminate(prob.forecast.demand[0], prob.forecast.prices[0])
    o.apg(3); o.extrapolate(0.4); o.solve_step()
    V, U_ref, X_ref = o.get("V").reshape(nodes, nv), o.get("U").reshape(nodes, nu), o.get("X").reshape(nodes, nx)
    uhat, e, uhat_prev = o.get("uhat").reshape(nodes, nu), o.get("e").reshape(nodes, nx), o.get("uhat_prev")
    L, B = _cm(o.get("L"), nu, nv), _cm(net.B, nx, nu)
    par = np.asarray(t.ancestor, dtype=int) - 1
    x_cur, u_prev = np.asarray(c.current_x, dtype=np.float64), np.asarray(c.prev_u, dtype=np.float64)
    U, X = np.zeros((nodes, nu)), np.zeros((nodes, nx))
    usum, esum = np.zeros((nodes, nu)), np.zeros((nodes, nx))
    for i in range(nodes):                       # breadth-first numbering: a parent precedes its children
        p = par[i]
        up, uhp = (u_prev, uhat_prev) if p < 0 else (U[p], uhat[p])
        U[i] = uhat[i] + (up - uhp) + L @ V[i]
        usum[i] = U[i] + (usum[p] if p >= 0 else 0.0)
        esum[i] = e[i] + (esum[p] if p >= 0 else 0.0)
        X[i] = x_cur + esum[i] + B @ usum[i]
    assert np.abs(U - U_ref).max() <= 1e-9 * np.abs(U_ref).max()
    assert np.abs(X - X_ref).max() <= 1e-9 * np.abs(X_ref).max()
    o.close()
