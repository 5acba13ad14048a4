#!/bin/bash
# GPU-box session: parity suite, then the default bench (4K, config4) and the reference arm.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log"
echo "== bench (default)"; timeout 900 python bench.py 2>"$OUT/bench.err" | tail -1 | tee "$OUT/bench_config4.json"
tail -5 "$OUT/bench.err"
