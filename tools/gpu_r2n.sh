#!/usr/bin/env bash
# 8-GPU diagnostic + validation: where does the partitioned solve differ from the 1-GPU solve (HEAD library), then the same
# check and the bench with the in-tree library.  Usage (under gpurun --gpus N): bash tools/gpu_r2n.sh <tag> <N>
set -uo pipefail
TAG="${1:-r2n}"; NG="${2:-8}"
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [ -f ab_libs/librapidnet_b200_head.so ]; then
  RAPIDNET_B200_LIB=$PWD/ab_libs/librapidnet_b200_head.so timeout 200 $TR --master-port 29511 tools/dist_check.py --workload C3 --iters 1,2 --diag > "$OUT/diag_head_C3_${NG}gpu.log" 2>&1
  echo "diag head rc=$?"; grep -E "it=|DIST_CHECK" "$OUT/diag_head_C3_${NG}gpu.log" | cut -c1-500
fi
timeout 300 $TR --master-port 29512 tools/dist_check.py --workload C3 --iters 1,10,100,500 --diag > "$OUT/dist_check_C3_${NG}gpu.log" 2>&1
echo "dist_check C3 x$NG rc=$?" | tee -a "$OUT/summary.txt"; grep -E "it=|DIST_CHECK" "$OUT/dist_check_C3_${NG}gpu.log" | cut -c1-500
timeout 300 $TR --master-port 29513 tools/dist_check.py --workload C3b --iters 1,100 --diag > "$OUT/dist_check_C3b_${NG}gpu.log" 2>&1
echo "dist_check C3b x$NG rc=$?" | tee -a "$OUT/summary.txt"; grep -E "it=|DIST_CHECK" "$OUT/dist_check_C3b_${NG}gpu.log" | cut -c1-500
timeout 600 $TR --master-port 29544 bench.py --gpus $NG --steps 5 --warmup 3 --closed-loop-instances 0 > "$OUT/bench_${NG}gpu.json" 2> "$OUT/bench_${NG}gpu.err"; echo "bench ${NG}gpu rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_${NG}gpu.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p=d.get("tree_partition",{})
    print("HEAD", d["n_gpus"], "gpus", d["scaling"], round(d["value"]), "iter/s  e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3))
    print("  check", p.get("check_vs_one_gpu"))
    print("  phases", p.get("iteration_ms_by_phase_rank0"), p.get("phase_clock_ns_per_iteration_rank0"))
    x=d.get("tree_partition_extra")
    if x: print("  extra", {k:x.get(k) for k in ("workload","value","ms_per_solve","error")}, x.get("check_vs_one_gpu"))
except Exception as ex:
    print("bench FAILED", ex)
PY
tail -3 "$OUT/bench_${NG}gpu.err"
