#!/bin/bash
# quick GPU check: parity suite + kernel-only bench at 1080p and 4K
set -u
cd "$(dirname "$0")/.."
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee "$OUT/pytest_gpu.log"
bash tests/quick_bench.sh --ni 1920 --nj 1080 --steps 3 --warmup 3 2>&1 | tee "$OUT/quick_1080p.log"
bash tests/quick_bench.sh --steps 3 --warmup 3 2>&1 | tee "$OUT/quick_4k.log"
