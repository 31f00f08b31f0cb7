#!/usr/bin/env bash
# A/B of the in-tree library against every library under ab_libs/ (tools/build_variants.sh): quick bench lines per workload,
# then the golden / parity tests with each variant.  Usage (under gpurun): bash tools/gpu_ab2.sh <tag> "<workloads>"
set -uo pipefail
TAG="${1:-ab2}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
QB="--steps 3 --warmup 3 --no-cpu-baseline --no-alt --by-config '' --closed-loop-instances 0"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], round(d["value"]), "iter/s frac", round(r["frac"],3), r["iteration_ms_by_phase"])
    ph=r["phase_clock_ns_per_iteration"]; print("   ", {k:v for k,v in ph.items() if not k.startswith("cyc.") and v})
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
for w in ${2:-C2}; do
  eval timeout 300 python bench.py $QB --workload $w > "$OUT/bench_tree_$w.json" 2> "$OUT/bench_tree_$w.err"; echo "bench in-tree $w rc=$?" >> "$OUT/summary.txt"
  show "$OUT/bench_tree_$w.json" "in-tree/$w"
  for lib in ab_libs/*.so; do
    [ -e "$lib" ] || continue
    n=$(basename "$lib" .so)
    eval RAPIDNET_B200_LIB="$PWD/$lib" timeout 300 python bench.py $QB --workload $w > "$OUT/bench_${n}_$w.json" 2> "$OUT/bench_${n}_$w.err"; echo "bench $n $w rc=$?" >> "$OUT/summary.txt"
    show "$OUT/bench_${n}_$w.json" "$n/$w"
  done
done
for lib in ab_libs/*.so; do
  [ -e "$lib" ] || continue
  n=$(basename "$lib" .so)
  RAPIDNET_B200_LIB="$PWD/$lib" timeout 300 python -m pytest tests/test_gpu_golden.py tests/test_gpu_edge_trees.py -m gpu -x -q --timeout 200 > "$OUT/pytest_$n.log" 2>&1; echo "pytest $n rc=$?" | tee -a "$OUT/summary.txt"
done
