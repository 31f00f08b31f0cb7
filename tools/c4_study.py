"""BASELINE config[3] at its stated size: closed-loop 24 h receding-horizon simulation of 1024 independent SMPC instances
(varied initial tank levels and demand / price forecasts, rapidnet_b200/closed_loop.py) on the shipped K = 30 tree, sharded over
the GPUs of the node -- replicas only, no data-path collective.  Every rank runs its instances as lanes of one GPU (several
factored handles side by side, each capped at a share of the SMs).  Run under torchrun, one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 \
      tools/c4_study.py --instances 1024 --steps 24 --factors full

Prints one JSON line (rank 0): solves/s of the whole job (wall clock around the study, max over ranks), APG iterations/s, the
closed-loop KPIs of SmpcController::updateKpi (/root/reference/src/SmpcController.cu:1778-1859) averaged over the instances,
and a checksum; --check N re-runs the first N instances of rank 0 on one handle and compares bit for bit."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def kpis(prob, controls, states, prices):
    """economic / smooth / safety / network sums of updateKpi for one instance (controls [steps, nu], states [steps+1, nx])"""
    n, c = prob.network, prob.config
    eco = smooth = safe = net = 0.0
    up = c.prev_u.astype(np.float64)
    for t in range(controls.shape[0]):
        u, x = controls[t].astype(np.float64), states[t + 1].astype(np.float64)
        eco += float(np.sum((n.alpha1.astype(np.float64) + prices[t][: n.nu].astype(np.float64)) * np.abs(u)))
        smooth += float(np.sum((up - u) ** 2))
        safe += float(np.sum(np.abs(np.minimum(x - n.xsafe.astype(np.float64), 0.0))))
        net += float(np.sum(np.abs(x)))
        up = u
    return eco, smooth, safe, net


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--workload", default="C1r30")
    ap.add_argument("--factors", default="full", choices=["full", "shared"])
    ap.add_argument("--check", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from rapidnet_b200 import cabi, closed_loop
    from rapidnet_b200.datagen import named_problem

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    prob = named_problem(args.workload, max_iter=args.iters)
    # the study needs `steps` forecast slots: the generator makes them per instance (closed_loop.make_instance)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    cap = max(1, sms // args.lanes)
    mode = cabi.FACTORS_FULL if args.factors == "full" else cabi.FACTORS_SHARED
    handles = []
    for _ in range(args.lanes):
        h = cabi.Solver(prob, device=local)
        h.set_stream(torch.cuda.Stream().cuda_stream)
        h.set_grid_limit(cap)
        h.set_modes(cabi.SWEEP_PERSISTENT, mode)
        h.factor_step()
        handles.append(h)
    closed_loop.simulate_lanes(handles, prob, args.lanes, 1, args.iters)            # warm-up: one solve per lane
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = closed_loop.simulate_lanes(handles, prob, args.instances, args.steps, args.iters, rank=rank, world=world)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t[0])
    # KPIs and a checksum of this rank's share
    acc = np.zeros(6)
    for idx, (us, xs) in res.items():
        inst = closed_loop.make_instance(prob, idx, args.steps)
        e, s, sa, ne = kpis(prob, us, xs, inst.prices)
        acc += np.array([e, s, sa, ne, float(np.isfinite(us).all() and np.isfinite(xs).all()), float(np.sum(us.astype(np.float64)))])
    if world > 1:
        t = torch.tensor(acc, dtype=torch.float64)
        dist.all_reduce(t)
        acc = t.numpy()
    ok = None
    if rank == 0 and args.check > 0:
        one = cabi.Solver(prob, device=local)
        one.set_grid_limit(cap)
        one.set_modes(cabi.SWEEP_PERSISTENT, mode)
        one.factor_step()
        ok = True
        for idx in list(closed_loop.shard(args.instances, world, 0))[: args.check]:
            us, xs = closed_loop.run_instance(one, prob, closed_loop.make_instance(prob, idx, args.steps), args.iters)
            ok = ok and np.array_equal(us, res[idx][0]) and np.array_equal(xs, res[idx][1])
        one.close()
    if rank == 0:
        solves = args.instances * args.steps
        print(json.dumps({"study": "C4: closed-loop Monte-Carlo", "workload": args.workload, "instances": args.instances,
                          "steps_per_instance": args.steps, "iterations_per_solve": args.iters, "n_gpus": world, "lanes_per_gpu": args.lanes,
                          "ctas_per_lane": cap, "factors": args.factors, "seconds": sec, "solves_per_s": solves / sec,
                          "apg_iterations_per_s": solves * args.iters / sec, "scaling": "weak (replicas only, no collective)",
                          "kpi_mean_per_instance": {"economic": acc[0] / args.instances, "smooth": acc[1] / args.instances,
                                                    "safety": acc[2] / args.instances, "network": acc[3] / args.instances},
                          "all_finite": bool(acc[4] == args.instances), "checksum_sum_of_controls": acc[5],
                          "lanes_equal_single_handle_bitwise": ok}), flush=True)
    for h in handles:
        h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
