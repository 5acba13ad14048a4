#!/bin/bash
# 8-GPU check of the default multi-rank bench path (--queue auto = the shared frame), kernel path only.
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/${2:-n8_auto}
mkdir -p "$OUT"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>"$OUT/bench_n${N}.err" | grep '^{' | tail -1 | tee "$OUT/bench_n${N}_auto.json"
tail -n 5 "$OUT/bench_n${N}.err"
