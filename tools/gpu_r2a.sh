#!/usr/bin/env bash
# Round-2 first GPU visit: data for decisions (FFMA2 A/B, C3 on one GPU with the L2 hints, L2 share on small/large trees,
# C5 on the fallback path, the two facade tests that were gated in round 1).  Usage (under gpurun): bash tools/gpu_r2a.sh <tag>
set -uo pipefail
TAG="${1:-r2a}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
QB="--steps 3 --warmup 3 --no-cpu-baseline --no-alt --closed-loop-instances 0"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]), "iter/s", d["roofline"]["iteration_ms_by_kernel"], "whole frac", round(d["roofline"]["whole_iteration"]["frac"],3))
    print("   phases", d["roofline"]["phase_clock_ns_per_iteration"])
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
RN_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_host_cpp.py -m gpu -x -q --timeout 200 > "$OUT/pytest_host.log" 2>&1; echo "pytest host rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_host.log"
for w in C2 C3; do
  timeout 300 python bench.py $QB --workload $w > "$OUT/bench_tree_$w.json" 2> "$OUT/bench_tree_$w.err"; echo "bench in-tree $w rc=$?" | tee -a "$OUT/summary.txt"
  show "$OUT/bench_tree_$w.json" "in-tree/$w"
  for lib in ab_libs/*.so; do
    [ -e "$lib" ] || continue
    n=$(basename "$lib" .so)
    RAPIDNET_B200_LIB="$PWD/$lib" timeout 300 python bench.py $QB --workload $w > "$OUT/bench_${n}_$w.json" 2> "$OUT/bench_${n}_$w.err"; echo "bench $n $w rc=$?" | tee -a "$OUT/summary.txt"
    show "$OUT/bench_${n}_$w.json" "$n/$w"
  done
done
for lib in ab_libs/*.so; do
  [ -e "$lib" ] || continue
  n=$(basename "$lib" .so)
  RAPIDNET_B200_LIB="$PWD/$lib" timeout 200 python -m pytest tests/test_gpu_golden.py tests/test_gpu_edge_trees.py -m gpu -x -q --timeout 150 > "$OUT/pytest_$n.log" 2>&1; echo "pytest $n rc=$?" | tee -a "$OUT/summary.txt"
done
for cfg in "C1r6 1.0" "C1r30 0.3" "C1r30 0.8" "C3 0.05" "C3 0.2"; do
  set -- $cfg
  RN_L2_KEEP=$2 timeout 200 python bench.py $QB --workload $1 > "$OUT/bench_l2_$1_$2.json" 2> "$OUT/bench_l2_$1_$2.err"; echo "bench l2 $1 $2 rc=$?" | tee -a "$OUT/summary.txt"
  show "$OUT/bench_l2_$1_$2.json" "l2keep=$2/$1"
done
for w in C1r6 C1r30 C1; do
  timeout 200 python bench.py $QB --workload $w > "$OUT/bench_tree_$w.json" 2> "$OUT/bench_tree_$w.err"; echo "bench in-tree $w rc=$?" | tee -a "$OUT/summary.txt"
  show "$OUT/bench_tree_$w.json" "in-tree/$w"
done
timeout 400 python bench.py --steps 2 --warmup 3 --iters 40 --no-cpu-baseline --no-alt --closed-loop-instances 0 --workload C5 > "$OUT/bench_C5.json" 2> "$OUT/bench_C5.err"; echo "bench C5 rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_C5.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("C5", d["value"], d["ms_per_step"], d["roofline"])
except Exception as ex:
    print("C5 FAILED", ex)
PY
tail -3 "$OUT/bench_C5.err"
