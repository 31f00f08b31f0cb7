"""Host side of the subtree partition (rapidnet_b200/partition.py) -- CPU only, world_size 2 over gloo.

The cut is aligned to the bottom-crown nodes (stage cs-1).  What crosses GPUs in the partitioned solve is (i) for every
bottom-crown node the sum of q and of r over its chain heads, formed by the one rank that owns them and stored into a
table indexed by crown node id on every rank, from which every rank finishes the replicated crown; (ii) the beta rows of
the bottom-crown nodes (once per solve; exact on the owner); (iii) the two squared prox distances; (iv) the
infeasibility log, merged on the host.  The kernels need a GPU; here the SAME index spaces and exchange rules are driven
with numpy stand-ins for the per-chain work, two gloo ranks, and checked against the undivided computation (and, for
beta, against the oracle)."""
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
    t = named_problem("C2").tree                    # [6, 5, 3]: 30 bottom-crown nodes with 3 chains each
    cs = chain_stage(t)
    parts = split_chains(t, cs, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 90 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert all(lo % 3 == 0 and hi % 3 == 0 for lo, hi in parts)           # cut only between bottom-crown nodes
    assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 3
    assert split_chains(named_problem("C3").tree, 3, 8) == [(60 * r, 60 * r + 60) for r in range(8)]
    with pytest.raises(ValueError):
        split_chains(named_problem("C1").tree, 1, 2)                       # one bottom-crown node (the root): nothing to cut
    with pytest.raises(ValueError):
        split_chains(t, cs, 31)


@pytest.mark.parametrize("name,world", [("C1r6", 2), ("C1r30", 3), ("C2", 8), ("C3", 8)])
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
        # aligned cut: a bottom-crown node has all of its chains on this rank or none
        lo_, hi_ = head_ranges(t, cs)
        b0, b1 = int(t.nodes_per_stage_cumul[cs - 1]), int(t.nodes_per_stage_cumul[cs])
        loc = lt.n_children[b0:b1]
        assert np.all((loc == 0) | (loc == (hi_ - lo_)[b0:b1]))
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
        # (i) heads: q_h = sum of c down the chain, for OWN chains only; S_p = sum of the heads of every OWNED bottom-crown
        # node in ascending child order; every rank receives every S row (table indexed by crown node id)
        c_loc = c_glob[m.local_to_global]
        kl, T = lp.tree.K, t.N - cs
        q_own = c_loc[n_crown:].reshape(T, kl, nx).sum(axis=0)
        par = t.ancestor.astype(int) - 1
        lt = lp.tree
        b0 = int(t.nodes_per_stage_cumul[cs - 1])
        mine = {}
        for p in range(b0, n_crown):
            nc = int(lt.n_children[p])
            if nc == 0:
                continue
            first = int(lt.n_children_cumul[p - 1]) + 1 if p > 0 else 1          # local id of the first child
            j0 = first - int(lt.nodes_per_stage_cumul[cs])                     # local chain index
            acc = q_own[j0].copy()
            for k in range(1, nc):
                acc = acc + q_own[j0 + k]                                       # fp32, ascending: the order of parent_sum
            mine[p] = acc
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        S = np.zeros((n_crown, nx), dtype=np.float32)
        owners = np.zeros(n_crown, dtype=int)
        for d_ in parts:
            for p, v in d_.items():
                S[p] = v; owners[p] += 1
        assert np.all(owners[b0:] == 1)                                         # every bottom-crown node has exactly one owner
        # every rank finishes the crown from the table: q_i = c_i + sum_{crown below} c + sum_{bottom p below i} S_p
        below = np.zeros((n_crown, nx))
        sb = S.astype(np.float64).copy()
        sb[:b0] = 0
        for j in range(n_crown - 1, 0, -1):
            below[par[j]] += c_glob[j] + below[j]
            sb[par[j]] += sb[j]
        q_crown = c_glob[:n_crown].astype(np.float64) + below + sb
        # undivided reference: subtree sums over the whole tree
        q_ref = c_glob.astype(np.float64).copy()
        for j in range(t.nodes - 1, 0, -1):
            q_ref[par[j]] += q_ref[j]
        assert np.allclose(q_crown, q_ref[:n_crown], rtol=1e-5, atol=1e-4)
        # the S rows do not depend on the number of ranks: the undivided tree gives the same bits
        q_all = c_glob[n_crown:].reshape(T, K, nx).sum(axis=0)
        hlo, hhi = head_ranges(t, cs)
        for p in mine:
            acc = q_all[hlo[p]].copy()
            for k in range(hlo[p] + 1, hhi[p]):
                acc = acc + q_all[k]
            assert np.array_equal(acc, mine[p])
        # (ii) beta of the bottom-crown nodes: the owner's row (all children local) equals the whole tree's, bit for bit
        o = Oracle(lp); o.L_.l.orc_null_space(o.h); o.factor_step(); o.update_state()
        og = Oracle(prob, L=o.get("L"), Lhat=o.get("Lhat")); og.factor_step(); og.update_state()
        og.eliminate(prob.forecast.demand[0], prob.forecast.prices[0]); o.eliminate(prob.forecast.demand[0], prob.forecast.prices[0])
        nv = prob.config.nv
        want = og.get("beta").reshape(-1, nv)
        got = o.get("beta").reshape(-1, nv)
        own = [p for p in range(b0, n_crown) if lt.n_children[p] > 0]
        assert np.array_equal(got[own], want[own])
        assert np.array_equal(got[:b0], want[:b0])                              # upper crown: children are crown nodes, replicated
        og.close(); o.close()
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
