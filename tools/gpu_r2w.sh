#!/usr/bin/env bash
# Round-2 quick validation of a kernel change on ONE GPU: the parity files that exercise the persistent kernel, then the bench line.
# Usage (under gpurun): bash tools/gpu_r2w.sh <tag> [pytest -k expression]
set -uo pipefail
TAG="${1:-r2w}"
KEXPR="${2:-}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
FILES="tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_shared_factors.py tests/test_gpu_edge_trees.py"
if [[ -n "$KEXPR" ]]; then
  timeout 1200 python -m pytest $FILES -m gpu -q --timeout 600 -rs -s -k "$KEXPR" > "$OUT/pytest_sel.log" 2>&1
else
  timeout 1200 python -m pytest $FILES -m gpu -q --timeout 600 -rs -s > "$OUT/pytest_sel.log" 2>&1
fi
echo "pytest sel rc=$?" | tee -a "$OUT/summary.txt"
grep -E "passed|failed|Error|worst" "$OUT/pytest_sel.log" | tail -40
timeout 900 python bench.py --steps 10 --warmup 5 ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("HEAD", round(d["value"]), "iter/s  e2e", round(d["e2e"]["value"]), " frac", round(d["roofline"]["frac"],3), " phaseS frac", round(d["roofline"]["phase_S"]["frac"],3))
    print("  phases", d["roofline"]["iteration_ms_by_phase"], d["roofline"]["phase_clock_ns_per_iteration"])
    for k,v in (d.get("by_config") or {}).items():
        print("  ", k, round(v.get("value",0),1), "frac", round(v.get("roofline",{}).get("frac",0),3), "e2e", round(v.get("e2e",{}).get("value",0)), v.get("error"))
    for k in ("alt_formulation","alt_formulation_shared"):
        if k in d: print("  ", k, round(d[k].get("value",0)))
except Exception as ex:
    print("bench FAILED", ex)
PY
tail -3 "$OUT/bench.err"
