#!/usr/bin/env bash
# Round-2 multi-GPU visit.  Usage (under gpurun --gpus N): bash tools/gpu_r2m.sh <tag> <N> [parts]   parts: "check bench nccl c3b"
set -uo pipefail
TAG="${1:-r2m}"
NG="${2:-2}"
PARTS="${3:-check bench nccl}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [[ " $PARTS " == *" check "* ]]; then
  for wl in ${CHECK_WORKLOADS:-C2 C3}; do
    timeout 600 $TR --master-port 29511 tools/dist_check.py --workload $wl --iters 1,10,100,500 > "$OUT/dist_check_${wl}_${NG}gpu.log" 2>&1; echo "dist_check $wl x$NG rc=$?" | tee -a "$OUT/summary.txt"
    grep -E "it=|DIST_CHECK" "$OUT/dist_check_${wl}_${NG}gpu.log"
  done
  if [ "$NG" = "2" ]; then
    timeout 600 $TR --master-port 29512 tools/dist_check.py --workload C2 --factors shared --iters 1,10,100 > "$OUT/dist_check_C2_shared_${NG}gpu.log" 2>&1; echo "dist_check C2 shared x$NG rc=$?" | tee -a "$OUT/summary.txt"
    grep -E "it=|DIST_CHECK" "$OUT/dist_check_C2_shared_${NG}gpu.log"
  fi
fi
if [[ " $PARTS " == *" bench "* ]]; then
  EXTRA="--partition-extra ''"
  [[ " $PARTS " == *" c3b "* ]] && EXTRA=""
  eval timeout 900 $TR --master-port 29544 bench.py --gpus $NG --steps 5 --warmup 3 --closed-loop-instances 0 $EXTRA > "$OUT/bench_${NG}gpu.json" 2> "$OUT/bench_${NG}gpu.err"; echo "bench ${NG}gpu rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_${NG}gpu.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p=d.get("tree_partition",{})
    print("HEAD", d["n_gpus"], "gpus", d["scaling"], round(d["value"]), "iter/s  e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3))
    print("  partition", {k:p.get(k) for k in ("value","ms_per_solve","exchange_bytes_per_iteration_per_rank","nodes_per_rank","error")})
    print("  check", p.get("check_vs_one_gpu"))
    print("  phases", p.get("iteration_ms_by_phase_rank0"), p.get("phase_clock_ns_per_iteration_rank0"))
    x=d.get("tree_partition_extra")
    if x: print("  extra", {k:x.get(k) for k in ("workload","value","ms_per_solve","error")}, x.get("check_vs_one_gpu"))
    r=d.get("replicas",{})
    print("  replicas", round(r.get("value",0)), "e2e", round(r.get("e2e",{}).get("value",0)))
except Exception as ex:
    print("bench FAILED", ex)
PY
  tail -3 "$OUT/bench_${NG}gpu.err"
fi
if [[ " $PARTS " == *" nccl "* ]]; then
  timeout 300 $TR --master-port 29546 tools/nccl_exchange_baseline.py --parents 80 --row 160 > "$OUT/nccl_baseline_${NG}gpu.log" 2>&1; echo "nccl baseline x$NG rc=$?" | tee -a "$OUT/summary.txt"
  grep NCCL_BASELINE "$OUT/nccl_baseline_${NG}gpu.log"
fi
if [[ " $PARTS " == *" c4 "* ]]; then
  for f in ${C4_FACTORS:-full shared}; do
    timeout 900 $TR --master-port 29550 tools/c4_study.py --instances ${C4_INSTANCES:-1024} --steps ${C4_STEPS:-24} --factors $f > "$OUT/c4_study_${f}_${NG}gpu.json" 2> "$OUT/c4_study_${f}_${NG}gpu.err"; echo "c4 study $f x$NG rc=$?" | tee -a "$OUT/summary.txt"
    grep '"study"' "$OUT/c4_study_${f}_${NG}gpu.json" | cut -c1-900
  done
fi
if [[ " $PARTS " == *" ref "* ]]; then
  timeout 900 $TR --master-port 29547 bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > "$OUT/bench_ref_${NG}gpu.json" 2> "$OUT/bench_ref_${NG}gpu.err"; echo "bench ref ${NG}gpu rc=$?" | tee -a "$OUT/summary.txt"
  cut -c1-300 "$OUT/bench_ref_${NG}gpu.json"
fi
