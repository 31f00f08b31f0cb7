#!/usr/bin/env bash
# A/B of library variants under ab_libs/ (tools/build_variants.sh): parity tests + bench (with the alt legs) for each, then the
# in-tree library's bench.  AB_WORKLOADS="C2 C3" adds workloads (default C2).
set -uo pipefail
TAG="${1:-ab}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
QB="--steps 5 --warmup 3 --no-cpu-baseline --closed-loop-instances 0"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "full", round(d["value"]), d["roofline"]["iteration_ms_by_kernel"])
for k in ("alt_formulation","alt_formulation_shared"):
    a=d.get(k)
    if a: print(sys.argv[2], k, round(a["value"]), a["roofline"].get("iteration_ms_by_kernel"))
PY
}
WL="${AB_WORKLOADS:-C2}"
for lib in ab_libs/*.so; do
  [ -e "$lib" ] || continue
  n=$(basename "$lib" .so)
  RAPIDNET_B200_LIB="$PWD/$lib" timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shared_factors.py tests/test_gpu_golden.py -m gpu -x -q --timeout 200 > "$OUT/pytest_$n.log" 2>&1; echo "pytest $n rc=$?" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/pytest_$n.log"
  for w in $WL; do
    RAPIDNET_B200_LIB="$PWD/$lib" timeout 300 python bench.py $QB --workload $w > "$OUT/bench_${n}_$w.json" 2> "$OUT/bench_${n}_$w.err"; echo "bench $n $w rc=$?" | tee -a "$OUT/summary.txt"
    show "$OUT/bench_${n}_$w.json" "$n/$w"
  done
done
for w in $WL; do
  timeout 300 python bench.py $QB --workload $w > "$OUT/bench_tree_$w.json" 2> "$OUT/bench_tree_$w.err"; echo "bench in-tree $w rc=$?" | tee -a "$OUT/summary.txt"
  show "$OUT/bench_tree_$w.json" "in-tree/$w"
done
