#!/usr/bin/env bash
# Multi-GPU visit: partitioned-tree tests (2 GPUs; streamed + shared factors), then the N-GPU bench line (weak instances + C3
# partition + closed-loop lanes), streamed and shared factors.  Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> [N=2]
set -uo pipefail
TAG="${1:-m2}"
NG="${2:-2}"
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpu.txt" 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 300 > "$OUT/pytest_multi.log" 2>&1; echo "pytest multi rc=$?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_multi.log"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $NG --steps 3 --warmup 3 --closed-loop-instances 2 --closed-loop-lanes 4 > "$OUT/bench_${NG}gpu.json" 2> "$OUT/bench_${NG}gpu.err"; echo "bench ${NG}gpu rc=$?" | tee -a "$OUT/summary.txt"
tail -c 3000 "$OUT/bench_${NG}gpu.json"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $NG --steps 3 --warmup 3 --factors shared --closed-loop-instances 0 > "$OUT/bench_${NG}gpu_shared.json" 2> "$OUT/bench_${NG}gpu_shared.err"; echo "bench ${NG}gpu shared rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_${NG}gpu_shared.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("shared", d["n_gpus"], "gpus", d["value"], d.get("tree_partition"))
PY
