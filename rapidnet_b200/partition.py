"""Subtree partition of ONE scenario tree across the GPUs of a node (BASELINE config[2]).

The reference is single-GPU; this is the host side of the B200 extension the north-star asks for.  The tree is cut at the
first stage of its non-branching tail ("chain stage" cs: below it every scenario is an independent chain):

  * the crown (stages < cs) is REPLICATED on every rank -- a few dozen nodes;
  * rank r owns a contiguous range of the K scenario chains (stages cs .. N-1), i.e. ~1/G of the nodes, of the Engine
    factor matrices and of the duals.

The cut is aligned to the nodes of the last crown stage (stage cs-1, the "bottom-crown" nodes): all chains below one of
them go to the same rank.  Each rank builds an ordinary, smaller problem (crown + its chains, nodes renumbered
breadth-first) and creates its handle on it; per APG iteration the ranks exchange -- inside the persistent kernel, over
NVLink peer memory mapped with CUDA IPC -- for every bottom-crown node they own the sum of G q and of r over its chain
heads (2 nv floats: what the reference's solveSumChildren leaves in the parent's slot, q already multiplied by G, so that
every rank can finish the crown) and the two squared prox distances.  Per solve, the beta rows of the bottom-crown nodes travel the same way.
Nothing else crosses GPUs.  torch.distributed hands the 64-byte IPC handles around and provides the host barrier.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .problem import Problem, Tree


def chain_stage(tree: Tree) -> int:
    """First stage of the non-branching tail: every node of a later stage is the only child of the node with the same
    index one stage up (same rule as the library, rapidnet_b200/csrc/rn_api.cu).  tree.N if there is no tail."""
    nps, cum = tree.nodes_per_stage, tree.nodes_per_stage_cumul
    par = tree.ancestor.astype(np.int64) - 1
    cs = tree.N - 1
    while cs > 0:
        if nps[cs] != nps[cs - 1]:
            break
        j = np.arange(nps[cs])
        if not np.array_equal(par[cum[cs] + j], cum[cs - 1] + j):
            break
        cs -= 1
    return int(cs)


def split_chains(tree: Tree, cs: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous chain ranges [lo, hi) per rank, cut only BETWEEN the bottom-crown nodes (stage cs-1) and balanced by the
    number of chains: the edge of rank r is the parent boundary closest to K r / world."""
    K = int(tree.nodes_per_stage[cs])
    if cs <= 0:
        raise ValueError("the tree has no crown stage to cut below")
    cum = tree.nodes_per_stage_cumul
    par = tree.ancestor.astype(np.int64) - 1
    heads_parent = par[cum[cs]: cum[cs] + K]
    # chain index where each bottom-crown node's heads start (children are contiguous per parent), plus the end
    bounds = np.concatenate([np.flatnonzero(np.diff(heads_parent, prepend=heads_parent[0] - 1)), [K]])
    if world < 1 or bounds.size - 1 < world:
        raise ValueError(f"cannot split {bounds.size - 1} bottom-crown subtrees over {world} ranks")
    edges = [0]
    for r in range(1, world):
        target = K * r / world
        cand = bounds[(bounds > edges[-1])]
        cand = cand[cand <= K - (world - r)] if cand.size else cand
        edges.append(int(cand[np.argmin(np.abs(cand - target))]))
    edges.append(K)
    if any(edges[r + 1] <= edges[r] for r in range(world)):
        raise ValueError(f"cannot split the chains over {world} ranks along the bottom-crown nodes")
    return [(edges[r], edges[r + 1]) for r in range(world)]


@dataclass
class PartitionMeta:
    world: int
    rank: int
    cs: int                      # chain stage
    n_crown: int                 # nodes of the replicated crown (same ids on every rank)
    K_global: int
    chain_lo: int                # this rank owns global chains [chain_lo, chain_hi)
    chain_hi: int
    head_lo: np.ndarray          # [n_crown] global chain-index range of the chain heads below each crown node
    head_hi: np.ndarray
    local_to_global: np.ndarray  # [local nodes] global node id of each local node


def head_ranges(tree: Tree, cs: int) -> Tuple[np.ndarray, np.ndarray]:
    """For every crown node: the range of chain indices (0..K-1, position inside stage cs) of the heads below it.
    Children are contiguous per parent (/root/reference/src/Utilities.cu:184-199), so it is one range."""
    cum = tree.nodes_per_stage_cumul
    n_crown = int(cum[cs])
    par = tree.ancestor.astype(np.int64) - 1
    lo = np.full(n_crown, np.iinfo(np.int32).max, dtype=np.int64)
    hi = np.zeros(n_crown, dtype=np.int64)
    for j in range(int(tree.nodes_per_stage[cs])):
        a = par[cum[cs] + j]
        while a >= 0:
            lo[a] = min(lo[a], j)
            hi[a] = max(hi[a], j + 1)
            a = par[a]
    lo[hi == 0] = 0
    return lo.astype(np.int32), hi.astype(np.int32)


def local_problem(problem: Problem, rank: int, world: int) -> Tuple[Problem, PartitionMeta]:
    """The problem rank `rank` of `world` solves: crown + its chains, renumbered stage by stage."""
    t = problem.tree
    cs = chain_stage(t)
    if cs >= t.N:
        raise ValueError("the tree has no non-branching tail to partition")
    cum, nps = t.nodes_per_stage_cumul.astype(np.int64), t.nodes_per_stage.astype(np.int64)
    K = int(nps[cs])
    if K != t.K:
        raise ValueError("the chains of the tail are not the scenarios of the tree")
    lo, hi = split_chains(t, cs, world)[rank]
    kl = hi - lo
    n_crown = int(cum[cs])
    # global ids of the local nodes, in local (breadth-first) order
    l2g = [np.arange(n_crown, dtype=np.int64)]
    for s in range(cs, t.N):
        l2g.append(cum[s] + np.arange(lo, hi))
    l2g = np.concatenate(l2g)
    g2l = -np.ones(t.nodes, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    nodes = int(l2g.size)
    stages = t.stages[l2g].astype(np.int32)
    lnps = np.concatenate([nps[:cs], np.full(t.N - cs, kl), [0]]).astype(np.int32)
    lcum = np.concatenate([[0], np.cumsum(lnps)]).astype(np.int32)
    par_g = t.ancestor.astype(np.int64) - 1
    anc = np.where(par_g[l2g] >= 0, g2l[np.maximum(par_g[l2g], 0)] + 1, 0).astype(np.int32)   # 1-based, root = 0
    n_nonleaf = nodes - kl
    # children of node i (local): its global children that are local; contiguous and in order by construction
    cnt = np.zeros(n_nonleaf, dtype=np.int32)
    for i in range(1, nodes):
        cnt[anc[i] - 1] += 1
    ccum = np.zeros(nodes, dtype=np.int32)
    ccum[:n_nonleaf] = np.cumsum(cnt)
    ccum[n_nonleaf:] = ccum[n_nonleaf - 1] if n_nonleaf > 0 else 0
    nd, nu = problem.network.nd, problem.network.nu
    ed = np.asarray(t.err_demand, dtype=np.float32).reshape(t.nodes, nd)[l2g].reshape(-1)
    ep = np.asarray(t.err_price, dtype=np.float32).reshape(t.nodes, nu)[l2g].reshape(-1)
    lt = Tree(N=t.N, K=kl, nodes=nodes, n_nonleaf=int(n_nonleaf), n_children_tot=nodes - 1, stages=stages,
              nodes_per_stage=lnps, nodes_per_stage_cumul=lcum,
              leaves=(lcum[t.N - 1] + np.arange(kl) + 1).astype(np.int32),
              children=np.arange(2, nodes + 1, dtype=np.int32), ancestor=anc, n_children=cnt, n_children_cumul=ccum,
              prob=np.asarray(t.prob, dtype=np.float32)[l2g], dim_demand=t.dim_demand, dim_price=t.dim_price,
              err_demand=ed, err_price=ep)
    lp = Problem(network=problem.network, tree=lt, config=copy.copy(problem.config), forecast=problem.forecast)
    hlo, hhi = head_ranges(t, cs)
    meta = PartitionMeta(world=world, rank=rank, cs=cs, n_crown=n_crown, K_global=K, chain_lo=lo, chain_hi=hi,
                         head_lo=hlo, head_hi=hhi, local_to_global=l2g)
    return lp, meta


def merge_pinf(parts: List[np.ndarray]) -> np.ndarray:
    """vecPrimalInfs (SmpcController.cu:1487-1495) from every rank's per-iteration (|res|, res) at its arg-max of the xi
    block and of the psi block: the global arg-max of each block, then the larger of the two signed values."""
    p = np.stack(parts)                       # [ranks, iters, 4]
    ix = np.argmax(p[:, :, 0], axis=0)
    ip = np.argmax(p[:, :, 2], axis=0)
    it = np.arange(p.shape[1])
    return np.maximum(p[ix, it, 1], p[ip, it, 3])


class DistributedSolver:
    """One tree on `world` GPUs, one process per GPU.  `group` is a torch.distributed process group (nccl or gloo:
    it only carries the IPC handles and the host-side gathers of results)."""

    def __init__(self, problem: Problem, rank: int, world: int, device: int, group=None):
        import torch.distributed as dist
        from . import cabi
        self.rank, self.world, self.group = rank, world, group
        self.global_problem = problem
        self.local, self.meta = local_problem(problem, rank, world)
        self.solver = cabi.Solver(self.local, device=device, chain_stage_hint=self.meta.cs)
        m = self.meta
        handle = self.solver.dist_prepare(world, rank, m.K_global, m.chain_lo, m.head_lo, m.head_hi)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        self.solver.dist_connect(handles)
        dist.barrier(group=group)

    def setup(self, slot: int = 0):
        import torch.distributed as dist
        c, fc = self.global_problem.config, self.global_problem.forecast
        s = self.solver
        s.factor_step()
        s.update_state(c.current_x, c.prev_u, c.prev_demand)
        self.eliminate_coupling(fc.demand[slot], fc.prices[slot])
        # the first solve also allocates the persistent kernel's buffers: do it once here, then line the ranks up, so that
        # no later launch is skewed by more than the in-kernel wait budget
        s.prepare_persistent()
        s.sync()
        dist.barrier(group=self.group)

    def eliminate_coupling(self, d_hat, alpha_hat):
        """Engine::eliminateInputDistubanceCoupling on the partition.  zeta_i = p_i dU_i - sum_children p_c dU_c
        (calculateZeta, /root/reference/src/Utilities.cu:100-131) of a bottom-crown node sums over its children, which all
        live on one rank (the cut is aligned to these nodes): that rank's beta row is exact and is pushed into every rank's
        staging table over NVLink; after the barrier every rank pulls the rows it does not own.  No host copy of data."""
        import torch.distributed as dist
        s, m = self.solver, self.meta
        s.eliminate_coupling(d_hat, alpha_hat)
        if self.world == 1 or m.cs == 0:
            return
        s.dist_sync_crown_beta(pull=False)
        s.sync()
        dist.barrier(group=self.group)
        s.dist_sync_crown_beta(pull=True)

    def apg_solve(self, iterations: int, want_u0=True, check=True):
        """Every rank must call this with the same iteration count (the in-kernel barriers span the GPUs).  `check`: read
        the timeout flag of the cross-GPU waits back (a device synchronisation)."""
        u0, _ = self.solver.apg_solve(iterations, want_u0=want_u0)
        if check and self.solver.dist_error():
            raise RuntimeError("a cross-GPU wait timed out: a peer rank did not reach the exchange")
        return u0

    def control_action(self, x, u_prev, d_prev, d_hat, alpha_hat, iterations: int):
        """SmpcController::controlAction(real_t*) on the partition: host buffers in on every rank, u0 (of the replicated
        root) out on every rank."""
        self.solver.update_state(x, u_prev, d_prev)
        self.eliminate_coupling(d_hat, alpha_hat)
        return self.apg_solve(iterations, want_u0=True, check=True)

    def exchange_bytes_per_iteration(self) -> int:
        """bytes this rank stores into its peers per APG iteration: S rows of its bottom-crown nodes + the prox distances"""
        t, m, n = self.local.tree, self.meta, self.local.network
        cum = t.nodes_per_stage_cumul
        owned = int(np.count_nonzero(t.n_children[int(cum[m.cs - 1]): int(cum[m.cs])])) if m.cs > 0 else 0
        return (owned * 2 * self.local.config.nv * 4 + 2 * 16 + 2 * 4) * (self.world - 1)

    def gather(self, name: str, dim: int) -> np.ndarray:
        """A per-node buffer ([nodes][dim]) assembled in GLOBAL node order on every rank."""
        import torch.distributed as dist
        loc = self.solver.read(name).reshape(-1, dim)
        parts = [None] * self.world
        dist.all_gather_object(parts, (self.meta.local_to_global, loc), group=self.group)
        out = np.zeros((self.global_problem.tree.nodes, dim), dtype=np.float32)
        for l2g, a in parts:
            out[l2g] = a          # crown rows arrive from every rank with identical values
        return out

    def close(self):
        self.solver.close()
