#!/usr/bin/env bash
# A/B of the L2 prefetch share (RN_L2_PREFETCH) and of the L2 keep share on several trees; quick bench lines only.
# Usage (under gpurun): bash tools/gpu_r2c.sh <tag> "<workloads>" "<prefetch shares>"
set -uo pipefail
TAG="${1:-r2c}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
QB="--steps 3 --warmup 3 --no-cpu-baseline --no-alt --by-config '' --closed-loop-instances 0"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], round(d["value"]), "iter/s frac", round(r["frac"],3), r["iteration_ms_by_phase"])
    ph=r["phase_clock_ns_per_iteration"]; print("   ", {k:v for k,v in ph.items() if not k.startswith("cyc.")})
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
for w in ${2:-C2}; do
  for pf in ${3:-0 0.2 0.4 0.6}; do
    eval RN_L2_PREFETCH=$pf timeout 300 python bench.py $QB --workload $w > "$OUT/bench_${w}_pf$pf.json" 2> "$OUT/bench_${w}_pf$pf.err"; echo "bench $w pf=$pf rc=$?" >> "$OUT/summary.txt"
    show "$OUT/bench_${w}_pf$pf.json" "$w pf=$pf"
  done
done
