"""One tree across several GPUs vs the same tree on one GPU (run under torchrun, one process per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_check.py --workload C1r30 --iters 1,10,100
Every rank solves the partitioned problem; rank 0 also solves the whole tree on its GPU and compares the gathered
iterates (expected: bit-identical -- the cut is aligned to the bottom-crown nodes, whose head sums and beta rows are
formed by the owning rank in the order of the single-GPU solve; the crown is computed from the same tables, the chains by
the same code).  Prints one line per iteration count and "DIST_CHECK OK" / "DIST_CHECK FAIL"."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C1r30")
    ap.add_argument("--iters", default="1,10,100")
    ap.add_argument("--factors", default="full")
    ap.add_argument("--bench", type=int, default=0, help="also time this many solves of 100 iterations")
    ap.add_argument("--diag", action="store_true", help="where do the iterates differ: per variable, stage and owning rank; crown copies rank by rank")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rapidnet_b200 import cabi
    from rapidnet_b200.datagen import named_problem
    from rapidnet_b200.partition import DistributedSolver, merge_pinf

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")           # carries only the IPC handles and the host-side gathers
    prob = named_problem(args.workload)
    ds = DistributedSolver(prob, rank, world, device=local)
    ds.solver.set_modes(cabi.SWEEP_PERSISTENT, {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[args.factors])
    ds.setup(slot=0)
    ref = None
    if rank == 0:
        ref = cabi.Solver(prob, device=local)
        ref.set_modes(cabi.SWEEP_PERSISTENT, {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[args.factors])
        ref.factor_step(); ref.update_state(); ref.eliminate_coupling(prob.forecast.demand[0], prob.forecast.prices[0])
    n = prob.network
    ok = True
    for iters in [int(x) for x in args.iters.split(",")]:
        dist.barrier()
        u0 = ds.apg_solve(iters)
        parts = [None] * world
        dist.all_gather_object(parts, ds.solver.pinf_parts(iters))
        got = {name: ds.gather(name, dim) for name, dim in (("VEC_U", n.nu), ("VEC_X", n.nx), ("VEC_UPDATE_XI", 2 * n.nx),
                                                            ("VEC_UPDATE_PSI", n.nu), ("VEC_DUAL_XI", 2 * n.nx))}
        if args.diag:
            m = ds.meta
            for name, dim in (("VEC_X", n.nx), ("VEC_U", n.nu), ("VEC_V", prob.config.nv), ("VEC_UPDATE_XI", 2 * n.nx), ("VEC_DUAL_XI", 2 * n.nx)):
                loc = ds.solver.read(name).reshape(-1, dim)[: m.n_crown]
                allc = [None] * world
                dist.all_gather_object(allc, loc)
                if rank == 0:
                    for r in range(1, world):
                        bad = np.flatnonzero((allc[r] != allc[0]).any(axis=1))
                        if bad.size:
                            print(f"  DIAG it={iters} {name}: crown copy of rank {r} differs from rank 0 at {bad.size} crown nodes, first {bad[:8]}, "
                                  f"max abs {float(np.abs(allc[r] - allc[0]).max()):.3e}", flush=True)
        if rank == 0:
            ru0, rinf = ref.apg_solve(iters, want_infs=True)
            if args.diag:
                stages = np.asarray(prob.tree.stages).reshape(-1)
                for name, a in got.items():
                    b = ref.read(name).reshape(a.shape)
                    bad = np.flatnonzero((a != b).any(axis=1))
                    if bad.size:
                        st, cnt = np.unique(stages[bad], return_counts=True)
                        rel = np.abs(a[bad].astype(np.float64) - b[bad]).max(axis=1) / np.maximum(np.abs(b[bad]).max(axis=1), 1e-30)
                        print(f"  DIAG it={iters} {name}: {bad.size} of {a.shape[0]} nodes differ from the 1-GPU solve; by stage {dict(zip(st.tolist(), cnt.tolist()))}; "
                              f"first nodes {bad[:10].tolist()}; worst row-relative diff {float(rel.max()):.3e} at node {int(bad[np.argmax(rel)])}", flush=True)
            worst = 0.0
            for name, a in got.items():
                b = ref.read(name).reshape(a.shape)
                err = float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
                worst = max(worst, err)
            pinf = merge_pinf(parts)
            pe = float(np.abs(pinf[:iters] - rinf[:iters]).max())          # every row, the last one (k_finalize) included
            same = all(np.array_equal(got[k], ref.read(k).reshape(got[k].shape)) for k in got)
            print(f"{args.workload} x{world} it={iters}: worst rel diff vs 1 GPU {worst:.2e}, bit-identical {same}, "
                  f"u0 diff {float(np.abs(u0 - ru0).max()):.2e}, pinf diff {pe:.2e}", flush=True)
            pe_ok = pe < 1e-3 * max(1.0, float(np.abs(rinf[: max(iters - 1, 1)]).max()))
            # every exchanged sum is formed by ONE rank in the order of the single-GPU solve: bit-identical is the expectation
            ok = ok and worst < 1e-6 and pe_ok
    if args.bench:
        dist.barrier()
        ds.apg_solve(100, want_u0=False); ds.solver.sync()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.bench):
            ds.apg_solve(100, want_u0=False)
        ds.solver.sync()
        dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            t1 = time.perf_counter()
            for _ in range(args.bench):
                ref.apg_solve(100, want_u0=False)
            ref.sync()
            d1 = time.perf_counter() - t1
            print(f"{args.workload}: {world} GPUs {args.bench * 100 / dt:.0f} iter/s, 1 GPU {args.bench * 100 / d1:.0f} iter/s", flush=True)
    if rank == 0:
        print("DIST_CHECK OK" if ok else "DIST_CHECK FAIL", flush=True)
    ds.close()
    if ref is not None:
        ref.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
