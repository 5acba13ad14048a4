#!/bin/bash
# Multi-GPU session: tests/gpu_multi.sh <N> [tag] -- the in-process multi-device test and bench.py under
# torchrun at 1..N ranks (both arms), exactly as the driver launches them.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L | tee "$OUT/gpus.txt"
echo "== multi-device tests"; timeout 600 python -m pytest tests -m gpu -q -k "multi_device or tiles_partition or trace_canvas" 2>&1 | tail -3 | tee "$OUT/pytest_multi.log"
n=1
while [ $n -le $N ]; do
  echo "== bench N=$n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee "$OUT/bench_n$n.json"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1 | tee "$OUT/bench_n$n.json"
  fi
  n=$((n*2))
done
echo "== reference arm under torchrun N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | grep '^{' | tail -1 | tee "$OUT/bench_reference_n$N.json"
