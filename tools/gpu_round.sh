#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, bench (both arms), launch list, one ncu --set full capture of the stream kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [quick]
set -uo pipefail
TAG="${1:-r01}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
timeout 1500 python -m pytest tests -m gpu ${PYTEST_X--x} -q -s --timeout 400 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "bench ref rc=$?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --steps 5 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench_ref.json" "$OUT/bench.json"
if [ "${2:-}" != "quick" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu-baseline --no-alt --closed-loop-instances 0 > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
  # the persistent kernel (one launch = `--iters` APG iterations) and, for the factor stream alone, the stand-alone k_stream of
  # the chain mode
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apg_persistent -s 3 -c 1 -o "$OUT/prof_persist" -f \
      python bench.py --steps 1 --warmup 1 --iters 10 --no-cpu-baseline --no-alt --closed-loop-instances 0 > "$OUT/ncu_full.log" 2>&1; echo "ncu full (persistent) rc=$?" | tee -a "$OUT/summary.txt"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 5 -c 1 -o "$OUT/prof_stream" -f \
      python bench.py --steps 1 --warmup 1 --iters 10 --sweep chain --no-cpu-baseline --no-alt --closed-loop-instances 0 > "$OUT/ncu_full_stream.log" 2>&1; echo "ncu full (k_stream) rc=$?" | tee -a "$OUT/summary.txt"
fi
