#!/usr/bin/env bash
# A/B of the L2 evict_last share of the factor stream: RN_L2_KEEP = share of the L2 capacity.
# Usage: bash tools/gpu_l2win.sh <tag> "<values>"   (L2_WORKLOAD=C3 for another workload).
set -uo pipefail
TAG="${1:-l2w}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persisting_l2_cache_max_size", None), "window max", getattr(p, "access_policy_max_window_size", None))
PY
for W in ${2:-0 0.5 0.75 1.0}; do
  RN_L2_KEEP=$W timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --closed-loop-instances 0 --workload "${L2_WORKLOAD:-C2}" > "$OUT/bench_w$W.json" 2> "$OUT/bench_w$W.err"; echo "bench w=$W rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench_w$W.json" "$W" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("w", sys.argv[2], "full", round(d["value"]), round(d["roofline"]["frac"],3), d["roofline"]["iteration_ms_by_kernel"])
a=d.get("alt_formulation")
if a: print("w", sys.argv[2], "df", round(a["value"]), round(a["roofline"]["frac"],3), a["roofline"]["iteration_ms_by_kernel"])
PY
done
