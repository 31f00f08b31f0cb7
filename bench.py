#!/usr/bin/env python
"""bench.py -- APG iterations/sec and ms per SMPC solve on the Barcelona DWN (BASELINE.json metric).

A "step" is one SMPC solve: `--iters` (default 500 = the reference's maxIterations) APG iterations on one
scenario tree (default workload C2: Barcelona-shaped DWN 63/114/88, N=24, tree [6,5,3] -> K=90 scenarios,
1927 nodes, 359 MB of Engine factor matrices -> larger than the 126 MB L2, so nothing is cache-resident between
iterations).  Synthetic data of the reference's shape (rapidnet_b200/datagen.py; the true Barcelona blobs are
missing from the reference).

  value : APG iterations/s, whole job, duals cold-started on the device, inputs resident in HBM; CUDA events on the
          launching stream, barrier + synchronize on both sides, max over ranks.
  e2e   : the same metric through rn_control_action (the reference's controlAction(real_t*)): HOST buffers in
          (state, previous control/demand, demand and price forecasts), affine-term refresh, the APG solve, u0 back
          to the host -- H2D and D2H inside the timed region.
  roofline : kernel level: SURVEY 8(d) Tier-A algorithmic bytes of one APG iteration over the CUDA-event time of one
          iteration of k_apg_persistent (the one kernel that runs) against MEASURED_PEAKS.json; the factor-stream phase
          (in-kernel %globaltimer clock of one CTA) is a sub-field.
  N = 1 : workload C2 (BASELINE config[1]); `by_config` carries the same three numbers (value, e2e, roofline fraction)
          for the other scenario counts of the metric ("vs scenario count"): C1, C1r6, C1r30, C3, C3b.
  N > 1 : the headline is STRONG scaling of ONE tree (BASELINE config[2]: C3, K=480, 10 171 nodes) cut below its last
          branching stage across the N GPUs, one process per GPU -- the crown replicated, the chains split, the near-root
          contributions and the prox distances exchanged inside the persistent kernel over NVLink peer memory
          (rapidnet_b200/partition.py); its e2e goes through the distributed controlAction, its iterates are checked
          against the same tree solved on one GPU in the same run.  N independent C2 instances (closed-loop Monte-Carlo
          replicas, no data-path collective) are the secondary record `replicas`.
  --impl reference : the reference's own CUDA/cuBLAS build (oracle/_ref/ref_driver, compiled from
          /root/reference/src in place) on the SAME workload as this arm's headline at that N (C2 at N=1, C3 at N>1; the
          reference is single-GPU) through its controlAction(real_t*); if that binary is missing, the CPU oracle port on
          the host cores.  Rank 0 only.  Its repetitions are capped so that the run ends within minutes.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "apg_iterations_per_sec"
UNIT = "iter/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w": float(np.median(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def rank_problem(workload, iters, rank):
    """The workload's problem; rank r > 0 is another Monte-Carlo instance (its own initial tank levels)."""
    from rapidnet_b200.datagen import named_problem
    prob = named_problem(workload, max_iter=iters)
    if rank > 0:
        rng = np.random.default_rng(977 + rank)
        c, n = prob.config, prob.network
        c.current_x = (rng.uniform(0.3, 0.9, size=n.nx) * n.xmax).astype(np.float32)
    return prob


def cpu_baseline(prob, iters_sample, threads):
    """The oracle port (oracle/rapidnet_oracle.c, OpenMP over nodes) on a bounded sample of the same workload."""
    from oracle.oracle import Oracle
    o = Oracle(prob, L=prob.config.L, Lhat=prob.config.Lhat, threads=threads)
    o.factor_step(); o.update_state(); o.eliminate(prob.forecast.demand[0], prob.forecast.prices[0])
    o.apg(2)                                   # warm the caches / thread pool
    per_solve = max(1, min(iters_sample, prob.config.max_iter))
    done, t0 = 0, time.perf_counter()
    while done < iters_sample:                 # whole cold-started solves, like the GPU arm's steps
        o.apg(per_solve); done += per_solve
    dt = time.perf_counter() - t0
    o.close()
    return done / dt, dt


def line_config(workload, prob, iters, desc=None):
    """`config` of the JSON line -- the same dict in both arms (the driver compares them)"""
    return {"workload": desc or describe(workload, prob), "iterations_per_solve": iters,
            "cold_cache": "inputs larger than L2: the factor matrices of one solve exceed the 126 MB L2"}


def headline_workload(args, world):
    """C2 on one GPU (BASELINE config[1]); the partitioned tree (config[2]) when the run spans several GPUs"""
    return args.workload if world == 1 or not args.partition_workload else args.partition_workload


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = headline_workload(args, world)
    prob = rank_problem(wl, args.iters, 0)
    # bounded: the reference needs seconds per solve (launch-bound on one host thread)
    reps, warm = max(1, min(args.steps, args.ref_max_steps)), min(args.warmup, 1)
    cfg = line_config(wl, prob, args.iters)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    have_gpu = shutil.which("nvidia-smi") is not None and subprocess.call(
        ["nvidia-smi", "-L"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "setup": {"repetitions": f"{reps} timed solves after {warm} warm-up (capped: --ref-max-steps)", "gpus_used": 1}}
    if world > 1:
        line["setup"]["note"] = "the reference is a single-GPU program: it solves the same tree on one GPU at every N"
    if os.path.exists(ref_bin) and have_gpu and not args.cpu_reference:
        from rapidnet_b200.problem import write_problem
        tmp = tempfile.mkdtemp(prefix="rn_ref_")
        cfg_path = write_problem(prob, tmp)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
        out = subprocess.run([ref_bin, cfg_path, "0", str(warm), str(reps)], capture_output=True,
                             text=True, env=env, timeout=3000)
        shutil.rmtree(tmp, ignore_errors=True)
        rec = [ln for ln in out.stdout.splitlines() if ln.startswith("REF ")]
        if out.returncode != 0 or not rec:
            line.update({"unavailable": f"ref_driver failed rc={out.returncode}: {(out.stderr or out.stdout)[-300:]}"})
            print(json.dumps(line)); return
        kv = dict(x.split("=") for x in rec[-1].split()[1:])
        ms = float(kv["ms_per_solve"])
        val = args.iters / (ms * 1e-3)
        line.update({"value": val, "ms_per_step": ms,
                     "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                                      "sample": "the reference's own CUDA/cuBLAS build (it has no CPU path): full "
                                                f"controlAction, {args.iters} iterations, median of {reps} on 1 GPU, "
                                                "one host thread driving cuBLAS"},
                     "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "reference_build": "oracle/_ref/ref_driver (unmodified /root/reference/src, nvcc sm_100, cuBLAS)"})
        print(json.dumps(line)); return
    # CPU oracle port on all host cores, bounded sample
    threads = os.cpu_count() or 1
    sample = max(4, args.cpu_sample_iters)
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, dt = cpu_baseline(prob, sample, threads)
        vals.append(v)
    val = float(np.median(vals))
    line.update({"value": val, "ms_per_step": args.iters / val * 1e3,
                 "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                  "sample": f"{sample} APG iterations (whole cold-started solves) of the same workload (oracle port, OpenMP)"},
                 "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def measure_handle(s, prob, iters, steps, warmup, stream, barrier, min_warmup=3):
    """One factored handle: `steps` cold-started solves with device-resident inputs (CUDA events on the launching stream),
    then the same through rn_control_action with host buffers.  Returns this rank's figures (ms over all steps)."""
    import torch
    c, fc = prob.config, prob.forecast
    with torch.cuda.stream(stream):
        for _ in range(max(warmup, min_warmup)):
            s.apg_solve(iters, want_u0=False)
        barrier()
        l0 = s.info().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            s.apg_solve(iters, want_u0=False)
        e1.record(stream)
        barrier()
        ms_dev = e0.elapsed_time(e1)
        launches = s.info().kernel_launches - l0
        host_in = [np.ascontiguousarray(a, dtype=np.float32) for a in
                   (c.current_x, c.prev_u, c.prev_demand, fc.demand[0], fc.prices[0])]
        u0 = np.zeros(prob.network.nu, dtype=np.float32)
        for _ in range(min(2, min_warmup)):
            s.control_action(*host_in, iters, out=u0)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            s.control_action(*host_in, iters, out=u0)
        f1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(f0.elapsed_time(f1), wall_ms)
    return {"ms_dev": ms_dev, "ms_e2e": ms_e2e, "launches": int(launches), "h2d": int(sum(a.nbytes for a in host_in)),
            "d2h": int(u0.nbytes), "finite": bool(np.isfinite(u0).all())}


def bench_config(name, args, local, stream, barrier, peak):
    """value / e2e / kernel-level roofline fraction of one scenario-count configuration on ONE GPU (by_config)"""
    from rapidnet_b200 import cabi
    prob = rank_problem(name, args.iters, 0)
    fc = prob.forecast
    s = cabi.Solver(prob, device=local)
    s.set_stream(stream.cuda_stream)
    s.set_modes({"persistent": cabi.SWEEP_PERSISTENT, "chain": cabi.SWEEP_CHAIN, "per_stage": cabi.SWEEP_PER_STAGE, "batched": cabi.SWEEP_BATCHED}[args.sweep],
                {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[args.factors])
    s.factor_step(); s.update_state(); s.eliminate_coupling(fc.demand[0], fc.prices[0])
    info = s.info()
    big = info.factor_bytes > 4e9            # seconds per solve: fewer repetitions
    steps = max(1, min(args.steps, 2 if big else 5))
    m = measure_handle(s, prob, args.iters, steps, 1 if big else 3, stream, barrier, min_warmup=1 if big else 3)
    info = s.info()
    it_s = m["ms_dev"] / steps / args.iters * 1e-3
    out = {"workload": describe(name, prob), "scenarios": int(prob.tree.K), "nodes": int(prob.tree.nodes), "steps": steps,
           "value": steps * args.iters / (m["ms_dev"] * 1e-3), "unit": UNIT, "ms_per_solve": m["ms_dev"] / steps,
           "e2e": {"value": steps * args.iters / (m["ms_e2e"] * 1e-3), "ms_per_solve": m["ms_e2e"] / steps},
           "persistent_kernel": bool(info.launches_per_iteration == 0),
           "launches_per_iteration": int(info.launches_per_iteration),
           "path": "k_apg_persistent" if info.launches_per_iteration == 0 else "k_stream + batched sweeps (GEMMs across all nodes) + k_finalize, CUDA graph per iteration",
           "factor_mb": info.factor_bytes / 1e6,
           "roofline": {"bytes_per_iteration": info.apg_bytes_per_iteration,
                        "achieved": info.apg_bytes_per_iteration / it_s / 1e9, "unit": "GB/s",
                        "frac": info.apg_bytes_per_iteration / it_s / 1e9 / peak,
                        "note": "factors fit the 126 MB L2: the HBM roofline does not bound this configuration"
                        if info.factor_bytes < 100e6 else None}}
    s.close()
    return out


def bench_partition(args, rank, world, local, stream, workload):
    """ONE tree cut across the `world` GPUs (strong scaling): iterations/s of the partitioned solve (device-resident inputs)
    and through the distributed controlAction (host buffers), max over ranks; rank 0 also solves the whole tree on its own
    GPU: the 1-GPU figure of the same workload and the iterate check of the partition."""
    import torch
    import torch.distributed as dist
    from rapidnet_b200 import cabi
    from rapidnet_b200.datagen import named_problem
    from rapidnet_b200.partition import DistributedSolver
    prob = named_problem(workload, max_iter=args.iters)
    c, fc, n = prob.config, prob.forecast, prob.network
    fmode = {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[args.factors]
    ds = DistributedSolver(prob, rank, world, device=local)
    ds.solver.set_stream(stream.cuda_stream)
    ds.solver.set_modes(cabi.SWEEP_PERSISTENT, fmode)
    ds.setup(slot=0)
    iters, steps = args.iters, max(1, min(args.steps, 5))
    host_in = [np.ascontiguousarray(a, dtype=np.float32) for a in (c.current_x, c.prev_u, c.prev_demand, fc.demand[0], fc.prices[0])]
    with torch.cuda.stream(stream):
        for _ in range(2):
            ds.apg_solve(iters, want_u0=False)
        dist.barrier(); torch.cuda.synchronize()
        l0 = ds.solver.info().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            ds.apg_solve(iters, want_u0=False, check=False)
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = ds.solver.info().kernel_launches - l0
        if ds.solver.dist_error():
            raise RuntimeError("a cross-GPU wait timed out during the timed solves")
        # end to end: host buffers in on every rank, u0 back on the host of every rank
        u0 = ds.control_action(*host_in, iters)
        dist.barrier(); torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            u0 = ds.control_action(*host_in, iters)
        f1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    # iterate check + the 1-GPU figure of the same tree (rank 0, the other GPUs idle)
    got = {name: ds.gather(name, dim) for name, dim in (("VEC_U", n.nu), ("VEC_X", n.nx), ("VEC_UPDATE_XI", 2 * n.nx))}
    check = None
    if rank == 0 and not args.no_partition_check:
        ref = cabi.Solver(prob, device=local)
        ref.set_stream(stream.cuda_stream)
        ref.set_modes(cabi.SWEEP_PERSISTENT, fmode)
        ref.factor_step()
        with torch.cuda.stream(stream):
            ru0 = ref.control_action(*host_in, iters)
            worst = 0.0
            for name, a in got.items():
                b = ref.read(name).reshape(a.shape).astype(np.float64)
                worst = max(worst, float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)))
            u0_err = float(np.linalg.norm(u0.astype(np.float64) - ru0) / max(np.linalg.norm(ru0), 1e-30))
            ref.apg_solve(iters, want_u0=False)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            one_steps = max(1, min(steps, 3))
            g0.record(stream)
            for _ in range(one_steps):
                ref.apg_solve(iters, want_u0=False)
            g1.record(stream)
            torch.cuda.synchronize()
            one_ms = g0.elapsed_time(g1) / one_steps
        check = {"iterations": iters, "worst_rel_err_U_X_y_vs_one_gpu": worst, "u0_rel_err_vs_one_gpu": u0_err,
                 "ok": bool(worst < 1e-4 and u0_err < 1e-4),
                 "one_gpu": {"value": iters / (one_ms * 1e-3), "unit": UNIT, "ms_per_solve": one_ms,
                             "note": "the same tree solved on ONE GPU of this box in the same run (rank 0)"}}
        ref.close()
    dist.barrier()
    # phase clock of the partitioned kernel (every rank runs the same solve: the in-kernel barriers span the GPUs)
    phases, prof = None, None
    try:
        with torch.cuda.stream(stream):
            prof = ds.solver.profile_kernels(min(iters, 100))
            phases = {k: round(v) for k, v in ds.solver.phase_times().items() if not k.startswith("cyc.")}
    except Exception as ex:   # noqa: BLE001
        phases = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    dist.barrier()
    d = prob.dims
    info = ds.solver.info()
    out = {"workload": describe(workload, prob), "n_gpus": world, "scaling": "strong", "steps": steps,
           "value": steps * iters / (ms * 1e-3), "unit": UNIT, "ms_per_solve": ms / steps,
           "e2e": {"value": steps * iters / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_solve": ms_e2e / steps,
                   "h2d_bytes_per_step": int(sum(a.nbytes for a in host_in)) * world, "d2h_bytes_per_step": int(u0.nbytes) * world,
                   "api": "DistributedSolver.control_action: rn_control_action on every rank (host buffers in, u0 out)"},
           "gpu_launches": int(launches),
           "nodes_per_rank": int(ds.local.tree.nodes), "crown_nodes_replicated": int(ds.meta.n_crown),
           "exchange_bytes_per_iteration_per_rank": int(ds.exchange_bytes_per_iteration()),
           "exchange": "in-kernel stores to peer memory (CUDA IPC over NVLink) + flag barriers; no NCCL on the data path",
           "bytes_per_iteration_per_rank": info.apg_bytes_per_iteration,
           "iteration_ms_by_phase_rank0": prof, "phase_clock_ns_per_iteration_rank0": phases,
           "check_vs_one_gpu": check}
    ds.close()
    return out


def bench_closed_loop(args, rank, world, local, stream, barrier):
    """A bounded sample of the 1024-instance x 24-step study: every rank factors once, then runs its instances' closed loops
    (rn_control_action with host buffers + rn_move_forward).  Weak scaling: instances per rank fixed."""
    import torch
    from rapidnet_b200 import cabi, closed_loop
    from rapidnet_b200.datagen import named_problem
    prob = named_problem(args.closed_loop_workload, max_iter=args.iters)
    s = cabi.Solver(prob, device=local)
    s.set_stream(stream.cuda_stream)
    s.factor_step()
    total = args.closed_loop_instances * world
    with torch.cuda.stream(stream):
        closed_loop.run_instance(s, prob, closed_loop.make_instance(prob, 0, 1), args.iters)   # warm-up
        barrier()
        t0 = time.perf_counter()
        res = closed_loop.simulate(s, prob, total, args.closed_loop_steps, args.iters, rank=rank, world=world)
        barrier()
        sec = time.perf_counter() - t0
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    solves = total * args.closed_loop_steps
    finite = all(np.isfinite(u).all() and np.isfinite(x).all() for u, x in res.values())
    out = {"workload": describe(args.closed_loop_workload, prob), "instances": total, "steps_per_instance": args.closed_loop_steps,
           "iterations_per_solve": args.iters, "scaling": "weak", "solves_per_s": solves / sec, "ms_per_closed_loop_step": sec / max(1, len(res) * args.closed_loop_steps) * 1e3,
           "apg_iterations_per_s": solves * args.iters / sec, "finite": bool(finite),
           "note": "instances shard with no data-path collective; one factored handle per GPU is reused by all of its instances"}
    s.close()
    # the same study with several handles of this GPU solving side by side (rn_set_grid_limit): a K-scenario tree keeps K
    # CTAs busy in its sweeps, the lanes fill the SMs one solve leaves idle.  Streamed and shared-factor formulations.
    lanes = args.closed_loop_lanes
    if lanes > 1:
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        cap = max(1, sms // lanes)
        per_rank = max(args.closed_loop_instances, 2 * lanes)
        handles = []
        for _ in range(lanes):
            h = cabi.Solver(prob, device=local)
            h.set_stream(torch.cuda.Stream().cuda_stream)
            h.set_grid_limit(cap)
            h.factor_step()
            handles.append(h)
        for key, mode in (("lanes", cabi.FACTORS_FULL), ("lanes_shared", cabi.FACTORS_SHARED)):
            for h in handles:
                h.set_modes(cabi.SWEEP_PERSISTENT, mode)
            closed_loop.simulate_lanes(handles, prob, lanes, 1, args.iters)   # warm-up: one solve per lane
            barrier()
            t0 = time.perf_counter()
            r2 = closed_loop.simulate_lanes(handles, prob, per_rank * world, args.closed_loop_steps, args.iters, rank=rank, world=world)
            barrier()
            sec2 = time.perf_counter() - t0
            t2 = torch.tensor([sec2], dtype=torch.float64, device="cuda")
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            sec2 = float(t2[0])
            n2 = per_rank * world * args.closed_loop_steps
            out[key] = {"lanes_per_gpu": lanes, "ctas_per_lane": cap, "factors": "full" if mode == cabi.FACTORS_FULL else "shared",
                        "instances": per_rank * world, "solves_per_s": n2 / sec2, "apg_iterations_per_s": n2 * args.iters / sec2,
                        "speedup_vs_one_handle": (n2 / sec2) / (solves / sec),
                        "finite": bool(all(np.isfinite(u).all() and np.isfinite(x).all() for u, x in r2.values()))}
        for h in handles:
            h.close()
    return out


def describe(workload, prob):
    d = prob.dims
    return (f"{workload}: Barcelona-shaped DWN nx={d['nx']} nu={d['nu']} nd={d['nd']} nv={d['nv']} N={d['N']}, "
            f"K={d['K']} scenarios, {d['nodes']} nodes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--iters", type=int, default=500, help="APG iterations per SMPC solve (reference: 500)")
    ap.add_argument("--cpu-sample-iters", type=int, default=2000, help="CPU baseline: APG iterations of the bounded sample (about 10-15 s on 16 cores)")
    ap.add_argument("--cpu-reference", action="store_true", help="--impl reference: force the CPU oracle port")
    ap.add_argument("--ref-max-steps", type=int, default=5, help="--impl reference: most timed solves (seconds each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the legs of the two reformulations (D, F only; shared factors)")
    ap.add_argument("--by-config", default="C1,C1r6,C1r30,C3,C3b,C5", help="N = 1: other scenario counts measured next to the headline ('' = skip)")
    ap.add_argument("--closed-loop-instances", type=int, default=4, help="closed-loop Monte-Carlo leg: instances PER RANK (0 = skip)")
    ap.add_argument("--closed-loop-lanes", type=int, default=4, help="that leg again with this many handles per GPU side by side (1 = skip)")
    ap.add_argument("--closed-loop-steps", type=int, default=2, help="receding-horizon steps per instance in that leg")
    ap.add_argument("--closed-loop-workload", default="C1r30", help="tree of that leg (SURVEY C4: the shipped K=30 tree)")
    ap.add_argument("--sweep", default="persistent", choices=["persistent", "chain", "per_stage", "batched"])
    ap.add_argument("--factors", default="full", choices=["full", "df", "shared"])
    ap.add_argument("--partition-workload", default="C3", help="N > 1: the tree that is cut across the GPUs = the headline ('' = replicas only)")
    ap.add_argument("--partition-extra", default="C3b", help="N > 1: a second, larger tree cut across the GPUs ('' = skip)")
    ap.add_argument("--no-partition-check", action="store_true", help="N > 1: skip the 1-GPU solve of the partitioned tree (check + 1-GPU figure)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from rapidnet_b200 import cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- rapidnet_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prob = rank_problem(args.workload, args.iters, rank)
    fc = prob.forecast
    s = cabi.Solver(prob, device=local)
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)          # torch.cuda.Event sees the stream the kernels are launched on
    s.set_modes({"persistent": cabi.SWEEP_PERSISTENT, "chain": cabi.SWEEP_CHAIN, "per_stage": cabi.SWEEP_PER_STAGE, "batched": cabi.SWEEP_BATCHED}[args.sweep],
                {"full": cabi.FACTORS_FULL, "df": cabi.FACTORS_DF, "shared": cabi.FACTORS_SHARED}[args.factors])
    s.factor_step()
    s.update_state()
    s.eliminate_coupling(fc.demand[0], fc.prices[0])
    iters = args.iters
    peak, peak_src = measured_peaks()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m = measure_handle(s, prob, iters, args.steps, args.warmup, stream, barrier)
    ms_dev, ms_e2e, launches = m["ms_dev"], m["ms_e2e"], m["launches"]
    clocks = sampler.stop() if rank == 0 and world == 1 else None

    # the same solve with the two exact reformulations of the factor step.  Reported next to the headline, never instead of
    # it; each roofline uses the bytes its own formulation has to move:
    #   df      only D, F streamed (v = -1/2 Omega r): half the matrix bytes
    #   shared  no per-node matrix at all ("Tier B"): D xi = G (sysF' xi), F psi = L' (s_u o psi) from the shared matrices
    alt, alt_shared = None, None
    if world == 1 and args.factors == "full" and args.sweep == "persistent" and not args.no_alt:
        def alt_leg(mode, text):
            s.set_modes(cabi.SWEEP_PERSISTENT, mode)
            with torch.cuda.stream(stream):
                s.apg_solve(iters, want_u0=False)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                for _ in range(args.steps):
                    s.apg_solve(iters, want_u0=False)
                g1.record(stream)
                barrier()
                ms_alt = g0.elapsed_time(g1)
            prof_alt = s.profile_kernels(min(iters, 100))
            info_alt = s.info()
            try:
                ph_alt = {k: round(v) for k, v in s.phase_times().items()}
            except Exception:
                ph_alt = None
            it_ms = ms_alt / args.steps / iters
            ach = info_alt.apg_bytes_per_iteration / (it_ms * 1e-3) / 1e9
            ach_s = info_alt.stream_bytes_per_iteration / (prof_alt["stream"] * 1e-3) / 1e9 if prof_alt["stream"] > 0 else 0.0
            return {"formulation": text, "value": args.steps * iters / (ms_alt * 1e-3), "unit": UNIT, "ms_per_solve": ms_alt / args.steps,
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                 "algorithmic_bytes_per_launch": info_alt.apg_bytes_per_iteration, "launch_ms": it_ms,
                                 "phase_S": {"bytes": info_alt.stream_bytes_per_iteration, "ms": prof_alt["stream"], "achieved": ach_s,
                                             "frac": ach_s / peak},
                                 "iteration_ms_by_phase": prof_alt, "phase_clock_ns_per_iteration": ph_alt}}
        def guarded(mode, text):
            try:
                return alt_leg(mode, text)
            except Exception as ex:   # noqa: BLE001  (e.g. a workload the persistent kernel does not fit)
                return {"formulation": text, "error": f"{type(ex).__name__}: {ex}"[:300]}
        alt = guarded(cabi.FACTORS_DF, "factors=df: only D, F streamed (v = -1/2 Omega r)")
        alt_shared = guarded(cabi.FACTORS_SHARED, "factors=shared (Tier B): no per-node matrix; D xi = G (sysF' xi), F psi = L' (s_u o psi) "
                                                  "from the shared matrices in shared memory, v = -1/2 Omega r")
        s.set_modes(cabi.SWEEP_PERSISTENT, cabi.FACTORS_FULL)

    # metric "vs scenario count": the other trees on ONE GPU, same three numbers each
    by_config = None
    if world == 1 and args.by_config:
        by_config = {}
        for name in [x for x in args.by_config.split(",") if x and x != args.workload]:
            try:
                by_config[name] = bench_config(name, args, local, stream, barrier, peak)
            except Exception as ex:   # noqa: BLE001
                by_config[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # closed-loop Monte-Carlo sample (BASELINE config[3]): instances sharded over the ranks, one factored handle per rank
    # the secondary legs never take the headline line down with them: an exception is recorded in their place
    loop = None
    if args.closed_loop_instances > 0:
        try:
            loop = bench_closed_loop(args, rank, world, local, stream, barrier)
        except Exception as ex:   # noqa: BLE001
            loop = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    part, part_extra = None, None
    if world > 1 and args.partition_workload:
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        try:
            part = bench_partition(args, rank, world, local, stream, args.partition_workload)
        except Exception as ex:   # noqa: BLE001
            part = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        clocks = sampler2.stop() if rank == 0 else None
        if args.partition_extra:
            try:
                part_extra = bench_partition(args, rank, world, local, stream, args.partition_extra)
            except Exception as ex:   # noqa: BLE001
                part_extra = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    elif world > 1 and rank == 0:
        clocks = sampler.stop()

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    value = world * args.steps * iters / (ms_dev * 1e-3)
    e2e_value = world * args.steps * iters / (ms_e2e * 1e-3)

    if rank == 0:
        info = s.info()
        prof = s.profile_kernels(min(iters, 100))
        try:
            phases = {k: round(v) for k, v in s.phase_times().items()} if args.sweep == "persistent" else None
        except Exception:
            phases = None
        it_s = ms_dev / args.steps / iters * 1e-3                      # CUDA-event time of one iteration of the one kernel that runs
        achieved = info.apg_bytes_per_iteration / it_s / 1e9
        ach_s = info.stream_bytes_per_iteration / (prof["stream"] * 1e-3) / 1e9 if prof["stream"] > 0 else 0.0
        total_prof = sum(prof.values())
        traffic, traffic_note = None, None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath) and args.workload == "C2" and args.sweep == "persistent" and args.factors == "full":
                tj = json.load(open(tpath))
                traffic = tj["dram_bytes_per_iteration"]
                traffic_note = (f"profiles/{tname}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of "
                                f"{tj['kernel']} ({tj['iterations_in_capture']} iterations) / iterations")
                break
        single = {
            "value": value, "unit": UNIT, "ms_per_solve": ms_dev / args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_solve": ms_e2e / args.steps,
                    "h2d_bytes_per_step": m["h2d"] * world, "d2h_bytes_per_step": m["d2h"] * world,
                    "api": "rn_control_action (SmpcController::controlAction(real_t*))"},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm",
                         "kernel": "k_apg_persistent: one APG iteration (factor stream + fused element-wise pass + tree sweeps + prox), "
                                   "CUDA events around the launch / iterations" if args.sweep == "persistent" else "per-iteration kernels",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_launch": info.apg_bytes_per_iteration, "launch_ms": it_s * 1e3,
                         "bytes_formula": "SURVEY.md 8(d) Tier A: per-node Engine factor matrices once + diagonals + vector traffic",
                         "phase_S": {"what": "factor-matrix stream + fused element-wise pass incl. its grid barrier (in-kernel %globaltimer clock of one CTA)",
                                     "bytes": info.stream_bytes_per_iteration, "ms": prof["stream"], "achieved": ach_s, "frac": ach_s / peak,
                                     "share_of_iteration": prof["stream"] / total_prof if total_prof > 0 else None},
                         "iteration_ms_by_phase": prof, "phase_clock_ns_per_iteration": phases},
        }
        setup = {"sweep": args.sweep, "factors": args.factors}
        if world > 1 and part is not None and "value" in part:
            # strong scaling of ONE tree is the headline of a multi-GPU run; the independent instances are secondary
            line = {
                "metric": METRIC, "value": part["value"], "unit": UNIT, "n_gpus": world, "steps": part["steps"],
                "warmup": max(args.warmup, 3), "ms_per_step": part["ms_per_solve"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": line_config(None, None, iters, desc=part["workload"]),
                "setup": dict(setup, parallelism=f"one tree cut across {world} GPUs below its last branching stage (crown replicated, chains split)"),
                "ms_per_solve": part["ms_per_solve"], "e2e": part["e2e"], "gpu_launches": part["gpu_launches"] * world,
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "k_apg_persistent on every rank, one APG iteration of the partitioned tree",
                             "achieved": part["bytes_per_iteration_per_rank"] * world / (part["ms_per_solve"] / iters * 1e-3) / 1e9,
                             "peak": peak * world, "unit": "GB/s",
                             "frac": part["bytes_per_iteration_per_rank"] / (part["ms_per_solve"] / iters * 1e-3) / 1e9 / peak,
                             "peak_source": peak_src + f" x {world} GPUs", "traffic": None,
                             "algorithmic_bytes_per_launch": part["bytes_per_iteration_per_rank"] * world,
                             "launch_ms": part["ms_per_solve"] / iters},
                "tree_partition": part,
                "replicas": dict(single, workload=describe(args.workload, prob), scaling="weak",
                                 parallelism=f"{world} independent SMPC instances, one per GPU, no data-path collective"),
            }
            if part_extra is not None:
                line["tree_partition_extra"] = part_extra
        else:
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": line_config(args.workload, prob, iters),
                "setup": dict(setup, instances=world, parallelism=f"{world} independent SMPC instance(s), one per GPU",
                              factor_mb=info.factor_bytes / 1e6),
                "ms_per_solve": single["ms_per_solve"], "e2e": single["e2e"], "gpu_launches": single["gpu_launches"],
                "launches_per_iteration": int(info.launches_per_iteration), "clocks": clocks, "roofline": single["roofline"],
            }
            if part is not None:
                line["tree_partition"] = part
        if by_config is not None:
            by_config[args.workload] = {"workload": describe(args.workload, prob), "scenarios": int(prob.tree.K), "nodes": int(prob.tree.nodes),
                                        "value": value, "unit": UNIT, "ms_per_solve": ms_dev / args.steps,
                                        "e2e": {"value": e2e_value, "ms_per_solve": ms_e2e / args.steps},
                                        "roofline": {"bytes_per_iteration": info.apg_bytes_per_iteration, "achieved": achieved,
                                                     "unit": "GB/s", "frac": achieved / peak}}
            line["by_config"] = by_config
        if alt is not None:
            line["alt_formulation"] = alt
        if alt_shared is not None:
            line["alt_formulation_shared"] = alt_shared
        if loop is not None:
            line["closed_loop"] = loop
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = max(4, args.cpu_sample_iters)
            v, dt = cpu_baseline(prob, sample, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{sample} APG iterations (whole cold-started solves) of the same workload ({dt:.1f} s, oracle port, OpenMP)"}
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
