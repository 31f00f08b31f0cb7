"""Closed-loop receding-horizon simulation of many independent SMPC instances (BASELINE config[3]).

The reference's main loop (`src/main.cu:27-63`) runs ONE controller in closed loop: controlAction -> moveForewardInTime
(`src/SmpcController.cu:1607-1717`).  A Monte-Carlo study repeats that for many instances that share the network, the
scenario tree and the controller configuration and differ only in their initial tank levels and forecasts
(x0 ~ U(0.3, 0.9) xmax, demand forecast = base x diurnal x (1 + 0.1 N(0,1)), prices x (1 + 0.1 N(0,1))).
Instances are independent: they shard over the GPUs with no data-path collective ("replicas only"), and on one GPU they
reuse one handle -- one factor step, one set of factor matrices in HBM -- because everything instance-specific enters
through the arguments of rn_control_action.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .problem import Problem


def shard(instances: int, world: int, rank: int) -> range:
    """Contiguous, balanced instance ids of `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    return range(instances * rank // world, instances * (rank + 1) // world)


@dataclass
class Instance:
    x0: np.ndarray            # initial tank levels [nx]
    demand: List[np.ndarray]  # per closed-loop step: demand forecast over the horizon [N*nd]
    prices: List[np.ndarray]  # per closed-loop step: price forecast over the horizon [N*nu]


def make_instance(prob: Problem, idx: int, steps: int, seed: int = 20240004) -> Instance:
    """Instance `idx` of the Monte-Carlo study (seeded by its index only: the same on every rank / GPU count)."""
    rng = np.random.default_rng([seed, idx])
    n, c, fc = prob.network, prob.config, prob.forecast
    N, nd, nu = fc.N, fc.dim_demand, fc.dim_prices
    x0 = (rng.uniform(0.3, 0.9, size=n.nx) * n.xmax).astype(np.float32)
    base_d = fc.demand[0].reshape(N, nd).astype(np.float64)
    base_p = fc.prices[0].reshape(N, nu).astype(np.float64)
    demand, prices = [], []
    for t in range(steps):
        d = np.roll(base_d, -t, axis=0) * (1.0 + 0.1 * rng.standard_normal((N, nd)))
        p = np.roll(base_p, -t, axis=0) * (1.0 + 0.1 * rng.standard_normal((N, nu)))
        demand.append(d.astype(np.float32).reshape(-1))
        prices.append(p.astype(np.float32).reshape(-1))
    return Instance(x0, demand, prices)


def run_instance(solver, prob: Problem, inst: Instance, iterations: int) -> Tuple[np.ndarray, np.ndarray]:
    """`len(inst.demand)` closed-loop steps of one instance on an already factored handle: controlAction with the clamp
    of the stream variant, then the plant update x+ = x + B u0 (quirks A.4-2 / A.4-3 of the reference are kept).
    Returns (applied controls [steps, nu], states [steps + 1, nx])."""
    c, nd = prob.config, prob.network.nd
    x, up, dp = inst.x0.copy(), c.prev_u.astype(np.float32).copy(), c.prev_demand.astype(np.float32).copy()
    us, xs = [], [x.copy()]
    for d_hat, a_hat in zip(inst.demand, inst.prices):
        solver.control_action(x, up, dp, d_hat, a_hat, iterations, clamp=True)
        x, up = solver.move_forward()
        dp = d_hat[:nd].astype(np.float32)
        us.append(up.copy()); xs.append(x.copy())
    return np.stack(us), np.stack(xs)


def simulate(solver, prob: Problem, instances: int, steps: int, iterations: int, rank: int = 0, world: int = 1):
    """This rank's share of the study; returns {instance id: (controls, states)}."""
    out = {}
    for idx in shard(instances, world, rank):
        out[idx] = run_instance(solver, prob, make_instance(prob, idx, steps), iterations)
    return out


def simulate_lanes(solvers, prob: Problem, instances: int, steps: int, iterations: int, rank: int = 0, world: int = 1):
    """The same study with several handles ("lanes") of ONE GPU solving side by side: lane l runs every len(solvers)-th
    instance of this rank's share on its own handle and CUDA stream, each handle's persistent kernel capped at a share
    of the SMs (`Solver.set_grid_limit`).  A tree with K scenarios keeps K CTAs busy in its sweeps, so the lanes fill SMs
    that one solve leaves idle.  One host thread per lane (the C-ABI calls block, ctypes releases the GIL).  Same
    result as `simulate`: instances are independent and every solve is deterministic for a given grid size."""
    from concurrent.futures import ThreadPoolExecutor
    mine = list(shard(instances, world, rank))
    lanes = len(solvers)

    def run_lane(l):
        return {idx: run_instance(solvers[l], prob, make_instance(prob, idx, steps), iterations) for idx in mine[l::lanes]}

    out = {}
    with ThreadPoolExecutor(max_workers=lanes) as ex:
        for part in ex.map(run_lane, range(lanes)):
            out.update(part)
    return out
