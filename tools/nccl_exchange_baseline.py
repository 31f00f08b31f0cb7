"""What the per-iteration exchange of the partitioned tree would cost with NCCL collectives (the plain design of SURVEY 8e:
build NCCL first, measure, then the in-kernel fast path): two ncclAllReduce per APG iteration -- the per-parent partial sums
[nodes(cs-1) x (nx + nv)] floats and the prox-distance / infeasibility scalars -- launched back to back from the host, timed
with CUDA events over many repetitions.  This is the communication floor of a host-driven design; the in-kernel exchange is
compared against it in DESIGN.md section 7.  Run under torchrun, one process per GPU."""
import argparse
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parents", type=int, default=80)
    ap.add_argument("--row", type=int, default=160, help="nx + nv")
    ap.add_argument("--reps", type=int, default=2000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    big = torch.zeros(args.parents * args.row, dtype=torch.float32, device="cuda")
    small = torch.zeros(4, dtype=torch.float64, device="cuda")
    for _ in range(50):
        dist.all_reduce(big); dist.all_reduce(small)
    torch.cuda.synchronize(); dist.barrier()
    out = {}
    for name, fn in (("partial_sums", lambda: dist.all_reduce(big)), ("scalars", lambda: dist.all_reduce(small)),
                     ("both_per_iteration", lambda: (dist.all_reduce(big), dist.all_reduce(small)))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.reps * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t[0])
    # the same inside a CUDA graph (no host launch cost between the collectives): 100 iterations' worth per launch
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            dist.all_reduce(big); dist.all_reduce(small)
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(g, stream=s):
                for _ in range(100):
                    dist.all_reduce(big); dist.all_reduce(small)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10):
                g.replay()
            e1.record(s)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 1000 * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["both_per_iteration_in_cuda_graph"] = float(t[0])
        except Exception as ex:  # noqa: BLE001
            out["both_per_iteration_in_cuda_graph"] = f"capture failed: {type(ex).__name__}"
    if rank == 0:
        print(f"NCCL_BASELINE world={world} payload={args.parents}x{args.row} floats: " +
              ", ".join(f"{k} {v:.1f} us" if isinstance(v, float) else f"{k} {v}" for k, v in out.items()), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
