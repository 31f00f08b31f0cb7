#!/usr/bin/env bash
# Short final check: core parity tests, then the default bench line.
set -uo pipefail
TAG="${1:-fin}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_shared_factors.py tests/test_gpu_reference.py -m gpu -x -q --timeout 150 > "$OUT/pytest_core.log" 2>&1; echo "pytest core rc=$?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_core.log"
timeout 300 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("full", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["roofline"]["iteration_ms_by_kernel"], "cpu", d.get("cpu_baseline"))
for k in ("alt_formulation","alt_formulation_shared"):
    print(k, round(d[k]["value"]))
print(d["closed_loop"])
PY
