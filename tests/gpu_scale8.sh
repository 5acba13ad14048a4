#!/bin/bash
# 8-GPU box: bench.py under torchrun at N = 4 and 8 (as the driver launches it) + in-process multi-device tests
set -u
cd "$(dirname "$0")/.."
TAG=${1:-scale8}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L | tee "$OUT/gpus.txt"
timeout 600 python -m pytest tests -m gpu -q -k "multi_device" 2>&1 | tail -3 | tee "$OUT/pytest_multi.log"
for n in 8 4; do
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 2>"$OUT/bench_n$n.err" | grep '^{' | tail -1 | tee "$OUT/bench_n$n.json"
done
