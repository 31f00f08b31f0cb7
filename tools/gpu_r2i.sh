#!/usr/bin/env bash
# Round-2 visit: facade / shim / closed-loop tests, the batched sweeps on large problems, then the full bench line.
set -uo pipefail
TAG="${1:-r2i}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_host_cpp.py tests/test_closed_loop.py tests/test_gpu_golden.py -m gpu -q --timeout 600 -rs > "$OUT/pytest_sel.log" 2>&1; echo "pytest selected rc=$?" | tee -a "$OUT/summary.txt"
tail -12 "$OUT/pytest_sel.log"
QB="--steps 2 --warmup 1 --no-cpu-baseline --no-alt --by-config '' --closed-loop-instances 0"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], round(d["value"],1), "iter/s frac", round(r["frac"],3), r["iteration_ms_by_phase"], "launches/iter", d.get("launches_per_iteration"))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
for cfg in "C3 batched" "C3 persistent" "C5 persistent" "C3b batched"; do
  set -- $cfg
  eval timeout 600 python bench.py $QB --workload $1 --sweep $2 > "$OUT/bench_$1_$2.json" 2> "$OUT/bench_$1_$2.err"; echo "bench $1 $2 rc=$?" >> "$OUT/summary.txt"
  show "$OUT/bench_$1_$2.json" "$1/$2"; tail -2 "$OUT/bench_$1_$2.err"
done
timeout 1200 python bench.py --steps 10 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("HEAD", round(d["value"]), "iter/s  e2e", round(d["e2e"]["value"]), " frac", round(d["roofline"]["frac"],3), " phaseS frac", round(d["roofline"]["phase_S"]["frac"],3))
    for k,v in (d.get("by_config") or {}).items():
        print("  ", k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_solve","error","launches_per_iteration")}, "frac", round(v.get("roofline",{}).get("frac",0),3), "e2e", round(v.get("e2e",{}).get("value",0),1))
    for k in ("alt_formulation","alt_formulation_shared"):
        if k in d: print("  ", k, round(d[k].get("value",0)))
    print("  closed_loop", {k:v for k,v in (d.get("closed_loop") or {}).items() if k in ("solves_per_s","error")}, "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as ex:
    print("bench FAILED", ex)
PY
tail -3 "$OUT/bench.err"
