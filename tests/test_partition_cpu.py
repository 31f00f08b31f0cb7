"""Host side of the subtree partition (rapidnet_b200/partition.py) -- CPU only, world_size 2 over gloo.

What crosses GPUs in the partitioned solve is (i) q and r of the chain heads, gathered into a table in global chain
order on every rank, from which every rank finishes the replicated crown; (ii) the zeta rows of the nodes just above
the heads (once per solve); (iii) the two squared prox distances; (iv) the infeasibility log, merged on the host.  The
kernels need a GPU; here the SAME index spaces and exchange rules are driven with numpy stand-ins for the per-chain
work, two gloo ranks, and checked against the undivided computation (and, for zeta, against the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle import Oracle
from rapidnet_b200.datagen import named_problem
from rapidnet_b200.partition import chain_stage, head_ranges, local_problem, merge_pinf, split_chains


def test_split_chains():
    assert split_chains(90, 8) == [(0, 11), (11, 22), (22, 33), (33, 45), (45, 56), (56, 67), (67, 78), (78, 90)]
    assert split_chains(2, 2) == [(0, 1), (1, 2)]
    with pytest.raises(ValueError):
        split_chains(2, 3)


@pytest.mark.parametrize("name,world", [("C1", 2), ("C1r6", 2), ("C1r30", 3), ("C2", 8)])
def test_local_problems_tile_the_tree(name, world):
    prob = named_problem(name)
    t = prob.tree
    cs = chain_stage(t)
    owned = np.zeros(t.nodes, dtype=int)
    for r in range(world):
        lp, m = local_problem(prob, r, world)
        lt = lp.tree
        # (a rank's tree can look non-branching further up -- e.g. one chain per rank: the library gets m.cs as a hint)
        assert m.cs == cs and chain_stage(lt) <= cs and lt.K == m.chain_hi - m.chain_lo
        assert np.array_equal(m.local_to_global[: m.n_crown], np.arange(m.n_crown))      # crown ids are shared
        owned[m.local_to_global[m.n_crown:]] += 1
        # local tree is a consistent reference-schema tree: parents precede children, children contiguous per parent
        par = lt.ancestor.astype(int) - 1
        assert np.all(par[1:] >= 0) and np.all(par[1:] < np.arange(1, lt.nodes)) and np.all(np.diff(par[1:]) >= 0)
        assert lt.nodes_per_stage_cumul[lt.N] == lt.nodes and lt.n_nonleaf == lt.nodes - lt.K
        assert np.array_equal(lt.stages, t.stages[m.local_to_global])
        assert np.array_equal(lt.prob, t.prob[m.local_to_global])
        # a parent's local children are the global ones that fall into this rank's chain range
        assert int(lt.n_children.sum()) == lt.nodes - 1
    assert np.all(owned[int(t.nodes_per_stage_cumul[cs]):] == 1)                          # every chain node has one owner
    lo, hi = head_ranges(t, cs)
    assert lo[0] == 0 and hi[0] == t.K                                                     # the root sees every head
    for i in range(int(t.nodes_per_stage_cumul[cs])):                                       # a node's range = union of its children's
        kids = np.nonzero(t.ancestor.astype(int) - 1 == i)[0]
        kids = kids[kids < t.nodes_per_stage_cumul[cs]]
        if kids.size:
            assert lo[i] == lo[kids].min() and hi[i] == hi[kids].max()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = named_problem(name)
        t = prob.tree
        lp, m = local_problem(prob, rank, world)
        cs, n_crown, K = m.cs, m.n_crown, m.K_global
        nx = prob.network.nx
        rng = np.random.default_rng(7)
        c_glob = rng.standard_normal((t.nodes, nx)).astype(np.float32)      # stand-in for c = sysF' xi_w of every node
        # (i) heads: q_h = sum of c down the chain, for OWN chains only; table in global chain order on every rank
        c_loc = c_glob[m.local_to_global]
        kl, T = lp.tree.K, t.N - cs
        q_own = c_loc[n_crown:].reshape(T, kl, nx).sum(axis=0)
        parts = [None] * world
        dist.all_gather_object(parts, (m.chain_lo, q_own))
        table = np.zeros((K, nx), dtype=np.float32)
        for lo_, q in parts:
            table[lo_: lo_ + q.shape[0]] = q
        # every rank finishes the crown from the table: q_i = c_i + sum_{crown below} c + sum_{heads in range} q_h
        par = t.ancestor.astype(int) - 1
        q_crown = np.zeros((n_crown, nx), dtype=np.float64)
        for i in range(n_crown - 1, -1, -1):
            q_crown[i] = c_glob[i] + table[m.head_lo[i]: m.head_hi[i]].astype(np.float64).sum(axis=0)
        below = np.zeros((n_crown, nx))
        for j in range(n_crown - 1, 0, -1):
            below[par[j]] += c_glob[j] + below[j]
        q_crown += below
        # undivided reference: subtree sums over the whole tree
        q_ref = c_glob.astype(np.float64).copy()
        for j in range(t.nodes - 1, 0, -1):
            q_ref[par[j]] += q_ref[j]
        assert np.allclose(q_crown, q_ref[:n_crown], rtol=1e-5, atol=1e-4)
        # (ii) zeta of the nodes just above the heads: sum_ranks zeta^r - (G-1) p dU, against the oracle on the whole tree
        o = Oracle(lp); o.L_.l.orc_null_space(o.h); o.factor_step(); o.update_state()
        nu = prob.network.nu
        uhat = None
        if cs > 0:
            og = Oracle(prob, L=o.get("L"), Lhat=o.get("Lhat")); og.factor_step(); og.update_state()
            og.eliminate(prob.forecast.demand[0], prob.forecast.prices[0]); o.eliminate(prob.forecast.demand[0], prob.forecast.prices[0])
            cum = t.nodes_per_stage_cumul
            n0, n1 = int(cum[cs - 1]), int(cum[cs])
            # the oracle keeps beta, not zeta: compare beta after the same correction expressed in beta space
            # beta = 2 (W L)' zeta + p L' alpha is linear in zeta, so  beta = sum_r beta^r - (G-1) beta(no children)
            uh = o.get("uhat").reshape(-1, nu)
            pr = lp.tree.prob
            up = np.where((par[n0:n1] >= 0)[:, None], uh[np.maximum(par[n0:n1], 0)], o.get("uhat_prev")[None, :] if cs == 1 else 0)
            Wv = o.get("Wv").reshape(-1, nu).T; L = o.get("L").reshape(-1, nu).T
            alpha = o.get("alpha").reshape(-1, nu)[n0:n1]
            pdu = pr[n0:n1, None] * (uh[n0:n1] - up)
            beta_nochild = 2.0 * pdu @ Wv + pr[n0:n1, None] * (alpha @ L)
            mine = o.get("beta").reshape(-1, prob.config.nv)[n0:n1]
            allb = [None] * world
            dist.all_gather_object(allb, mine)
            fixed = sum(b.astype(np.float64) for b in allb) - (world - 1) * beta_nochild
            want = og.get("beta").reshape(-1, prob.config.nv)[n0:n1]
            assert np.linalg.norm(fixed - want) / np.linalg.norm(want) < 1e-5
            og.close()
        o.close()
        # (iii) distances: every rank's share, crown counted by rank 0 only, summed in rank order
        d_node = rng.random(t.nodes)
        share = d_node[m.local_to_global[n_crown:]].sum() + (d_node[:n_crown].sum() if rank == 0 else 0.0)
        shares = [None] * world
        dist.all_gather_object(shares, share)
        assert abs(sum(shares) - d_node.sum()) < 1e-9 * t.nodes
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["C1r6", "C1r30"])
def test_exchange_rules_two_ranks_gloo(name):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() == "ok"


def test_merge_pinf():
    # two ranks, three iterations: (|res|, res) of the xi block, then of the psi block
    a = np.array([[5, -5, 1, 1], [2, 2, 9, -9], [1, 1, 1, 1]], dtype=np.float32)
    b = np.array([[4, 4, 3, -3], [7, -7, 2, 2], [1, -1, 6, 6]], dtype=np.float32)
    # it0: xi arg-max on rank a (-5), psi on rank b (-3) -> max(-5, -3); it1: max(-7, -9); it2: xi tie -> first rank (1), psi 6
    assert np.array_equal(merge_pinf([a, b]), np.array([-3, -7, 6], dtype=np.float32))
