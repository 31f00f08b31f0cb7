#!/usr/bin/env bash
# ncu evidence for the shared-factor formulation: launch list + one --set full capture of k_apg_persistent (10 iterations).
set -uo pipefail
TAG="${1:-ncu_sh}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
B="--steps 1 --warmup 1 --iters 10 --factors shared --no-cpu-baseline --no-alt --closed-loop-instances 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --iters 20 --factors shared --no-cpu-baseline --no-alt --closed-loop-instances 0 > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_apg_persistent -s 3 -c 1 -o "$OUT/prof_persist_shared" -f python bench.py $B > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT"
