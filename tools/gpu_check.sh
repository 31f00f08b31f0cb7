#!/usr/bin/env bash
# Lean GPU visit: parity tests of the default (persistent) path first, a quick bench of the in-tree library and of every
# A/B library under ab_libs/, then the rest of the GPU suite.  Usage (under gpurun): bash tools/gpu_check.sh <tag>
set -uo pipefail
TAG="${1:-chk}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
QB="--steps 5 --warmup 3 --no-cpu-baseline --no-alt --closed-loop-instances 0"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q --timeout 300 > "$OUT/pytest_core.log" 2>&1; echo "pytest core rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_core.log"
timeout 200 python bench.py $QB > "$OUT/bench_quick.json" 2> "$OUT/bench_quick.err"; echo "bench quick rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_quick.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("HEAD", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["phase_clock_ns_per_iteration"])
PY
for lib in ab_libs/*.so; do
  [ -e "$lib" ] || continue
  n=$(basename "$lib" .so)
  RAPIDNET_B200_LIB="$PWD/$lib" timeout 200 python bench.py $QB > "$OUT/bench_$n.json" 2> "$OUT/bench_$n.err"; echo "bench $n rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_$n.json" "$n" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["frac"])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_golden.py > "$OUT/pytest_rest.log" 2>&1; echo "pytest rest rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_rest.log"
