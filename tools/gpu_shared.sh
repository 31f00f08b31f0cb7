#!/usr/bin/env bash
# GPU visit for the shared-factor formulation: canary (toy, short timeout), its parity tests, bench with the alt legs,
# default-path A/B against every library under ab_libs/.  Usage (under gpurun): bash tools/gpu_shared.sh <tag>
set -uo pipefail
TAG="${1:-sh}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 90 python -m pytest tests/test_gpu_shared_factors.py -m gpu -x -q -s --timeout 60 -k "toy or need" > "$OUT/pytest_canary.log" 2>&1; rc=$?
echo "canary rc=$rc" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/pytest_canary.log"
if [ $rc -eq 0 ]; then
  timeout 300 python -m pytest tests/test_gpu_shared_factors.py -m gpu -q -s --timeout 200 -k "not toy and not need" > "$OUT/pytest_shared.log" 2>&1; echo "pytest shared rc=$?" | tee -a "$OUT/summary.txt"
  tail -25 "$OUT/pytest_shared.log"
  timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --closed-loop-instances 0 > "$OUT/bench_alt.json" 2> "$OUT/bench_alt.err"; echo "bench alt rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_alt.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("full", d["value"], d["roofline"]["frac"], d["roofline"]["iteration_ms_by_kernel"])
for k in ("alt_formulation","alt_formulation_shared"):
    a=d.get(k)
    if a: print(k, a["value"], a["ms_per_solve"], a["roofline"]["frac"], a["roofline"].get("iteration_ms_by_kernel"), a["roofline"].get("phase_clock_ns_per_iteration"))
PY
fi
for lib in ab_libs/*.so; do
  [ -e "$lib" ] || continue
  n=$(basename "$lib" .so)
  RAPIDNET_B200_LIB="$PWD/$lib" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-alt --closed-loop-instances 0 > "$OUT/bench_$n.json" 2> "$OUT/bench_$n.err"; echo "bench $n rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_$n.json" "$n" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["frac"])
PY
done
