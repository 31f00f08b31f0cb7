"""Phase clock of the partitioned persistent kernel on EVERY rank (one process per GPU, under torchrun): where the time of an
iteration goes per rank and per clock CTA (RN_CLOCK_CTA).  python -m torch.distributed.run ... tools/dist_phases.py --workload C3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rapidnet_b200 import cabi
    from rapidnet_b200.datagen import named_problem
    from rapidnet_b200.partition import DistributedSolver
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    prob = named_problem(args.workload)
    ds = DistributedSolver(prob, rank, world, device=local)
    ds.solver.set_modes(cabi.SWEEP_PERSISTENT, cabi.FACTORS_FULL)
    ds.setup(slot=0)
    ds.apg_solve(args.iters, want_u0=False)
    dist.barrier()
    prof = ds.solver.profile_kernels(args.iters)
    ph = {k: round(v) for k, v in ds.solver.phase_times().items() if not k.startswith("cyc.") and v}
    outs = [None] * world
    dist.all_gather_object(outs, (rank, prof, ph))
    if rank == 0:
        for r, p, h in sorted(outs):
            print(f"PHASES {args.workload} x{world} clock_cta={os.environ.get('RN_CLOCK_CTA', '0')} rank {r}: " +
                  ", ".join(f"{k} {v * 1e3:.1f}us" for k, v in p.items()) + f" | {h}", flush=True)
    ds.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
