"""Seeded generators of benchmark inputs in the reference's JSON schema.

The true Barcelona topology and forecasts are missing blobs in the reference
(/root/reference/.MISSING_LARGE_BLOBS:1-10).  What ships --
src/paser/dataSource/controllerConfig32.json (L, Lhat, W, preconditioner, state,
prices) and two real trees -- is carried in data/barcelona_base.npz; B, Gd, E, Ed,
bounds and the demand forecast are synthesised here with a fixed seed so that the
oracle, the reference build and the CUDA path all read identical files.
RNG: numpy.random.default_rng(PCG64), seed = 20240001 + config index.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np

from .problem import Config, Forecast, Network, Problem, Tree

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_NPZ = os.path.join(_ROOT, "data", "barcelona_base.npz")
SEED0 = 20240001

# named tree shapes of BASELINE.md section 3
TREES = {
    "C1": [2],            # K=2,    47 nodes
    "C2": [6, 5, 3],      # K=90,   1927 nodes
    "C3": [10, 8, 6],     # K=480,  10171 nodes
    "C3b": [10, 8, 6, 4],  # K=1920, 38971 nodes
    "C5": [8, 8, 6],      # K=384,  8137 nodes (with the 4x network)
    "C5s": [2, 2],        # K=4,    91 nodes (4x network, small tree: parity tests at the 4x dimensions)
}


def make_tree(branching: Sequence[int], N: int, nd: int, nu: int, seed: int,
              demand_sigma: float = 70.0, price_sigma: float = 2.2, zero_frac: float = 0.35) -> Tree:
    """Breadth-first numbered scenario tree, children contiguous per parent."""
    rng = np.random.default_rng(seed)
    nps = [1]
    for s in range(1, N):
        b = branching[s - 1] if s - 1 < len(branching) else 1
        nps.append(nps[-1] * b)
    cum = np.concatenate([[0], np.cumsum(nps)]).astype(np.int64)
    nodes = int(cum[-1])
    stages = np.zeros(nodes, dtype=np.int32)
    ancestor = np.zeros(nodes, dtype=np.int32)
    prob = np.ones(nodes, dtype=np.float64)
    n_children = []
    for s in range(N):
        stages[cum[s]:cum[s + 1]] = s
        if s == N - 1:
            break
        b = branching[s] if s < len(branching) else 1
        for j in range(nps[s]):
            par = cum[s] + j
            kids = cum[s + 1] + j * b + np.arange(b)
            ancestor[kids] = par + 1
            w = rng.uniform(0.05, 1.0, size=b)
            prob[kids] = prob[par] * w / w.sum()
            n_children.append(b)
    n_nonleaf = int(cum[N - 1])
    n_children = np.asarray(n_children, dtype=np.int32)
    ncc = np.zeros(nodes, dtype=np.int32)
    ncc[:n_nonleaf] = np.cumsum(n_children)
    ncc[n_nonleaf:] = nodes - 1
    K = nps[-1]
    err_d = rng.normal(0.0, demand_sigma, size=(nodes, nd)) * (rng.uniform(size=(nodes, nd)) >= zero_frac)
    err_p = rng.normal(0.0, price_sigma, size=(nodes, nu)) * (rng.uniform(size=(nodes, nu)) >= zero_frac)
    err_d[0] = 0.0
    err_p[0] = 0.0
    return Tree(
        N=N, K=K, nodes=nodes, n_nonleaf=n_nonleaf, n_children_tot=nodes - 1, stages=stages,
        nodes_per_stage=np.asarray(nps + [0], dtype=np.int32),
        nodes_per_stage_cumul=np.concatenate([cum, [nodes]]).astype(np.int32),
        leaves=(np.arange(cum[N - 1], nodes) + 1).astype(np.int32),
        children=(np.arange(1, nodes) + 1).astype(np.int32), ancestor=ancestor,
        n_children=n_children, n_children_cumul=ncc, prob=prob.astype(np.float32),
        dim_demand=nd, dim_price=nu, err_demand=err_d.astype(np.float32).reshape(-1),
        err_price=err_p.astype(np.float32).reshape(-1))


def real_tree(tag: str) -> Tree:
    """One of the two Barcelona-size trees that ship with the reference ('32': K=6, '65': K=30)."""
    z = np.load(BASE_NPZ)
    kw = {}
    for name, f in Tree.__dataclass_fields__.items():
        v = z[f"tree{tag}.{name}"]
        kw[name] = int(v) if f.type in ("int", int) else np.ascontiguousarray(v)
    return Tree(**kw)


def _diurnal(hours: np.ndarray) -> np.ndarray:
    return 1.0 + 0.35 * np.sin(2 * np.pi * (hours - 7.0) / 24.0) + 0.15 * np.sin(4 * np.pi * (hours - 3.0) / 24.0)


def barcelona_problem(tree: Tree, seed: int = SEED0, sim_horizon: int = 4, max_iter: int = 500) -> Problem:
    """'Barcelona-shaped' network (63/114/88/17, nv=97) around the shipped controllerConfig32 pieces."""
    z = np.load(BASE_NPZ)
    rng = np.random.default_rng(seed)
    nx, nu, nd, ne, nv, N = 63, 114, 88, 17, 97, 24
    assert tree.N == N and tree.dim_demand == nd and tree.dim_price == nu
    L = z["cfg.matL"].astype(np.float64).reshape(nv, nu).T            # nu x nv
    Lhat = z["cfg.matLhat"].astype(np.float64).reshape(nd, nu).T      # nu x nd
    # E: orthonormal basis of the orthogonal complement of range(L)  => E L = 0; Ed = -E Lhat
    Uf, _, _ = np.linalg.svd(L, full_matrices=True)
    E = Uf[:, nv:].T                                                  # ne x nu
    Ed = -E @ Lhat                                                    # ne x nd
    # B: entries {-1,0,+1}, <= 2 non-zeros per column, every row non-empty
    B = np.zeros((nx, nu))
    rows = np.concatenate([rng.permutation(nx), rng.integers(0, nx, size=nu - nx)])
    for c in range(nu):
        B[rows[c], c] = 1.0
        if rng.uniform() < 0.6:
            r2 = int(rng.integers(0, nx))
            if r2 != rows[c]:
                B[r2, c] = -1.0
    Gd = np.zeros((nx, nd))
    Gd[rng.integers(0, nx, size=nd), np.arange(nd)] = -1.0
    cur_x = z["cfg.currentX"].astype(np.float64)
    # the shipped prevU (+-1.6e6) belongs to the missing true topology; rebuild a consistent one
    # from the shipped prevV and prevDemand: u = Lhat d + L v  (SmpcController.cu:676-693)
    prev_u = Lhat @ z["cfg.prevDemand"].astype(np.float64) + L @ z["cfg.prevV"].astype(np.float64)
    xmax = 1.2 * np.maximum(cur_x, 1000.0)
    xsafe = 0.35 * xmax
    umax = 1.5 * np.abs(prev_u) + 100.0
    umin = np.where(prev_u < 0, -umax, 0.0)
    net = Network(nx=nx, nu=nu, nd=nd, ne=ne, A=np.eye(nx, dtype=np.float32).reshape(-1),
                  B=B.T.reshape(-1).astype(np.float32), Gd=Gd.T.reshape(-1).astype(np.float32),
                  E=E.T.reshape(-1).astype(np.float32), Ed=Ed.T.reshape(-1).astype(np.float32),
                  xmin=np.zeros(nx, np.float32), xmax=xmax.astype(np.float32),
                  xsafe=xsafe.astype(np.float32), umin=umin.astype(np.float32),
                  umax=umax.astype(np.float32), alpha1=z["cfg.costAlpha1"].astype(np.float32), N=N)
    cfg = Config(nx=nx, nu=nu, nd=nd, nv=nv, N=N, L=z["cfg.matL"].astype(np.float32),
                 Lhat=z["cfg.matLhat"].astype(np.float32), costW=z["cfg.costW"].astype(np.float32),
                 penalty_x=float(z["cfg.penaltyStateX"][0]), penalty_xs=float(z["cfg.penaltySafetyX"][0]),
                 precond=z["cfg.matDiagPrecnd"].astype(np.float32), current_x=cur_x.astype(np.float32),
                 prev_u=prev_u.astype(np.float32), prev_demand=z["cfg.prevDemand"].astype(np.float32),
                 step_size=float(z["cfg.stepSize"][0]), max_iter=max_iter, ne=ne)
    base_d = z["cfg.prevDemand"].astype(np.float64)
    alpha2 = z["cfg.costAlpha2"].astype(np.float64).reshape(N, nu)
    fc = Forecast(N=N, sim_horizon=sim_horizon, dim_demand=nd, dim_prices=nu)
    for t in range(sim_horizon):
        hrs = (t + 1 + np.arange(N)) % 24
        fc.demand.append((base_d[None, :] * _diurnal(hrs)[:, None]).astype(np.float32).reshape(-1))
        fc.prices.append(np.roll(alpha2, -t, axis=0).astype(np.float32).reshape(-1))
    return Problem(net, tree, cfg, fc)


def scaled_problem(tree: Tree, scale: int = 4, seed: int = SEED0 + 5, sim_horizon: int = 2,
                   max_iter: int = 500) -> Problem:
    """Fully synthetic DWN with `scale` x the Barcelona dimensions (config C5); keeps nu > nx."""
    rng = np.random.default_rng(seed)
    nx, nu, nd, ne, N = 63 * scale, 114 * scale, 88 * scale, 17 * scale, 24
    nv = nu - ne
    assert tree.N == N and tree.dim_demand == nd and tree.dim_price == nu
    E = np.zeros((ne, nu))
    for r in range(ne):                      # mixing-node balance rows: a few +-1 per row, full row rank
        cols = rng.choice(nu, size=4, replace=False)
        E[r, cols] = rng.choice([-1.0, 1.0], size=4)
        E[r, r] = 1.0
    Ed = np.zeros((ne, nd))
    Ed[np.arange(ne), rng.integers(0, nd, size=ne)] = -1.0
    Q, _ = np.linalg.qr(E.T, mode="complete")
    L = Q[:, ne:]
    Lhat = -np.linalg.pinv(E) @ Ed
    B = np.zeros((nx, nu))
    rows = np.concatenate([rng.permutation(nx), rng.integers(0, nx, size=nu - nx)])
    for c in range(nu):
        B[rows[c], c] = 1.0
        if rng.uniform() < 0.6:
            r2 = int(rng.integers(0, nx))
            if r2 != rows[c]:
                B[r2, c] = -1.0
    Gd = np.zeros((nx, nd))
    Gd[rng.integers(0, nx, size=nd), np.arange(nd)] = -1.0
    xmax = rng.uniform(1000.0, 30000.0, size=nx)
    cur_x = rng.uniform(0.4, 0.8, size=nx) * xmax
    prev_d = rng.uniform(5.0, 300.0, size=nd)
    prev_u = rng.uniform(10.0, 800.0, size=nu)
    umax = 3.0 * prev_u + 100.0
    precond = np.tile(np.concatenate([rng.uniform(0.5, 1.5, nu), rng.uniform(0.01, 0.1, nx),
                                      rng.uniform(0.01, 0.1, nx)]), N)
    net = Network(nx=nx, nu=nu, nd=nd, ne=ne, A=np.eye(nx, dtype=np.float32).reshape(-1),
                  B=B.T.reshape(-1).astype(np.float32), Gd=Gd.T.reshape(-1).astype(np.float32),
                  E=E.T.reshape(-1).astype(np.float32), Ed=Ed.T.reshape(-1).astype(np.float32),
                  xmin=np.zeros(nx, np.float32), xmax=xmax.astype(np.float32),
                  xsafe=(0.35 * xmax).astype(np.float32), umin=np.zeros(nu, np.float32),
                  umax=umax.astype(np.float32), alpha1=rng.uniform(0, 0.15, nu).astype(np.float32), N=N)
    cfg = Config(nx=nx, nu=nu, nd=nd, nv=nv, N=N, L=L.T.reshape(-1).astype(np.float32),
                 Lhat=Lhat.T.reshape(-1).astype(np.float32),
                 costW=np.eye(nu, dtype=np.float32).reshape(-1), penalty_x=1e10, penalty_xs=1e7,
                 precond=precond.astype(np.float32), current_x=cur_x.astype(np.float32),
                 prev_u=prev_u.astype(np.float32), prev_demand=prev_d.astype(np.float32),
                 step_size=1e-4, max_iter=max_iter, ne=ne)
    fc = Forecast(N=N, sim_horizon=sim_horizon, dim_demand=nd, dim_prices=nu)
    for t in range(sim_horizon):
        hrs = (t + 1 + np.arange(N)) % 24
        fc.demand.append((prev_d[None, :] * _diurnal(hrs)[:, None]).astype(np.float32).reshape(-1))
        fc.prices.append(rng.uniform(0.0, 0.25, size=(N, nu)).astype(np.float32).reshape(-1))
    return Problem(net, tree, cfg, fc)


def named_problem(name: str, max_iter: int = 500) -> Problem:
    """C1, C1r6, C1r30, C2, C3, C3b (Barcelona-shaped) or C5 / C5s (4x synthetic) of BASELINE.md section 3."""
    idx = {"C1": 1, "C1r6": 11, "C1r30": 12, "C2": 2, "C3": 3, "C3b": 4, "C5": 5, "C5s": 15}[name]
    seed = SEED0 + idx
    if name == "C1r6":
        return barcelona_problem(real_tree("32"), seed=seed, max_iter=max_iter)
    if name == "C1r30":
        return barcelona_problem(real_tree("65"), seed=seed, max_iter=max_iter)
    if name in ("C5", "C5s"):
        tree = make_tree(TREES[name], 24, 88 * 4, 114 * 4, seed)
        return scaled_problem(tree, 4, seed=seed, max_iter=max_iter)
    tree = make_tree(TREES[name], 24, 88, 114, seed)
    return barcelona_problem(tree, seed=seed, max_iter=max_iter)
