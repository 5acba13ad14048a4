#!/bin/bash
# 1-GPU session for the shared frame (cross-GPU tile queue): the whole parity suite incl. tests/test_gpu_frame.py
# (two PROCESSES on one GPU through CUDA IPC), smoke, the default bench, and the same bench drawing from the
# shared queue (system-scope atomics, stores into the frame) to show that the shared path costs nothing at N = 1.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-frame}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
echo "== frame tests, verbose"; timeout 300 python -m pytest tests/test_gpu_frame.py -m gpu -q -s 2>&1 | tail -12 | tee "$OUT/pytest_frame.log"
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as e; e.smoke()' 2>&1 | tail -3 | tee "$OUT/smoke.log"
echo "== bench --queue shared (kernel only)"; timeout 600 python bench.py --queue shared --no-cpu-baseline --no-e2e 2>"$OUT/bench_shared.err" | tail -1 | tee "$OUT/bench_config4_shared_queue.json"
echo "== bench (default)"; timeout 900 python bench.py 2>"$OUT/bench.err" | tail -1 | tee "$OUT/bench_config4.json"
tail -5 "$OUT/bench_shared.err" "$OUT/bench.err"
ls -la "$OUT"
