#!/usr/bin/env bash
# GPU visit: several handles per GPU side by side (rn_set_grid_limit).  Usage (under gpurun): bash tools/gpu_lanes.sh <tag>
set -uo pipefail
TAG="${1:-ln}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_closed_loop.py -m gpu -x -q -s --timeout 150 > "$OUT/pytest_lanes.log" 2>&1; echo "pytest lanes rc=$?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_lanes.log"
for L in 2 4 6; do
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --closed-loop-instances 8 --closed-loop-lanes $L > "$OUT/bench_lanes$L.json" 2> "$OUT/bench_lanes$L.err"; echo "bench lanes $L rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_lanes$L.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
c=d["closed_loop"]
print("one handle", c["solves_per_s"], c["ms_per_closed_loop_step"])
for k in ("lanes","lanes_shared"):
    if k in c: print(k, c[k])
PY
done
