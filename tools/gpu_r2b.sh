#!/usr/bin/env bash
# Round-2 validation visit on ONE GPU: the whole GPU suite, the parity table, the bench line, the sanitizer logs.
# Usage (under gpurun): bash tools/gpu_r2b.sh <tag> [parts]   parts: any of "tests table bench sanit ncu" (default: all but ncu)
set -uo pipefail
TAG="${1:-r2b}"
PARTS="${2:-tests table bench sanit}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1; nproc >> "$OUT/gpu.txt"
if [[ " $PARTS " == *" tests "* ]]; then
  RN_RUN_UNVALIDATED=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rs > "$OUT/pytest_gpu.log" 2>&1; echo "pytest gpu rc=$?" | tee -a "$OUT/summary.txt"
  tail -15 "$OUT/pytest_gpu.log"
fi
if [[ " $PARTS " == *" table "* ]]; then
  timeout 1500 python tools/parity_table.py --configs "${TABLE_CONFIGS:-toy,C1,C1r6,C1r30,C2,C3}" --iters 1,10,100,500 --out "$OUT/parity_table" > "$OUT/parity_table.log" 2>&1; echo "parity table rc=$?" | tee -a "$OUT/summary.txt"
  cat "$OUT/parity_table.log" | tail -40
fi
if [[ " $PARTS " == *" bench "* ]]; then
  timeout 900 python bench.py --steps 10 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
  python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("HEAD", round(d["value"]), "iter/s  e2e", round(d["e2e"]["value"]), " frac", round(d["roofline"]["frac"],3), " phaseS frac", round(d["roofline"]["phase_S"]["frac"],3))
    print("  phases", d["roofline"]["phase_clock_ns_per_iteration"])
    for k,v in (d.get("by_config") or {}).items():
        print("  ", k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_solve","error")}, "frac", round(v.get("roofline",{}).get("frac",0),3), "e2e", round(v.get("e2e",{}).get("value",0)))
    for k in ("alt_formulation","alt_formulation_shared"):
        if k in d: print("  ", k, round(d[k].get("value",0)))
    print("  closed_loop", {k:v for k,v in (d.get("closed_loop") or {}).items() if k in ("solves_per_s","error")}, "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as ex:
    print("bench FAILED", ex)
PY
  tail -3 "$OUT/bench.err"
  timeout 900 python bench.py --impl reference --steps 10 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"
  cut -c1-400 "$OUT/bench_ref.json"
fi
if [[ " $PARTS " == *" sanit "* ]]; then
  for tool in memcheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool --log-file "$OUT/sanitizer_$tool.log" python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitizer_$tool.out" 2>&1; echo "sanitizer $tool rc=$?" | tee -a "$OUT/summary.txt"
    tail -3 "$OUT/sanitizer_$tool.log"
  done
  # the Barcelona-size path (TMA ring, 16 warps, crown + parent sums) under memcheck: C1r6, 3 iterations
  DBG_ITERS=1,3 timeout 900 compute-sanitizer --tool memcheck --log-file "$OUT/sanitizer_memcheck_C1r6.log" python tools/dbg_persist.py C1r6 > "$OUT/sanitizer_memcheck_C1r6.out" 2>&1; echo "sanitizer memcheck C1r6 rc=$?" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/sanitizer_memcheck_C1r6.log"
fi
if [[ " $PARTS " == *" ncu "* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_C2.csv" python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu-baseline --no-alt --by-config "" --closed-loop-instances 0 > "$OUT/ncu_launch.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
  ncu --set full --clock-control none --import-source on -k regex:k_apg_persistent -c 1 -o "$OUT/ncu_full_k_apg_persistent_C2" -f python bench.py --steps 1 --warmup 3 --iters 10 --no-cpu-baseline --no-alt --by-config "" --closed-loop-instances 0 > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?" | tee -a "$OUT/summary.txt"
fi
if [[ " $PARTS " == *" c4 "* ]]; then
  for f in shared full; do
    timeout 300 python tools/c4_study.py --instances ${C4_INSTANCES:-32} --steps ${C4_STEPS:-4} --factors $f > "$OUT/c4_study_${f}_1gpu.json" 2> "$OUT/c4_study_${f}_1gpu.err"; echo "c4 study $f rc=$?" | tee -a "$OUT/summary.txt"
    grep '"study"' "$OUT/c4_study_${f}_1gpu.json" | cut -c1-600
  done
fi
