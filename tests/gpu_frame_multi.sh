#!/bin/bash
# N-GPU session for the shared frame: tests/gpu_frame_multi.sh <N> [tag] -- the multi-device tests, then bench.py
# under torchrun at N ranks with the static deal and with the shared queue (kernel path only).
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-frame_multi}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi -L | tee "$OUT/gpus.txt"
nvidia-smi topo -m 2>/dev/null | head -12 | tee "$OUT/topo.txt"
echo "== multi-device tests"; timeout 300 python -m pytest tests/test_gpu_frame.py tests/test_gpu_parity.py -m gpu -q -rs -k "frame or multi_device" 2>&1 | tail -6 | tee "$OUT/pytest_multi.log"
for q in shared static; do
  echo "== bench N=$N --queue $q"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 8 --warmup 3 --queue $q --no-e2e --no-cpu-baseline 2>"$OUT/bench_n${N}_$q.err" | grep '^{' | tail -1 | tee "$OUT/bench_n${N}_$q.json"
done
tail -3 "$OUT"/*.err
