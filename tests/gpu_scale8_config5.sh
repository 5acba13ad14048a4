#!/bin/bash
# 8-GPU box: BASELINE configs[4] -- Kerr-Schild a=0.9, 7680x4320, tolerance sweep, tiles sharded over 8 ranks
set -u
cd "$(dirname "$0")/.."
TAG=${1:-scale8c5}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for tol in 1e-6 1e-8 1e-10; do
  echo "== config5 tol=$tol N=8"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus 8 --workload config5 --tol $tol --steps 3 --warmup 3 --no-e2e 2>"$OUT/c5_$tol.err" | grep '^{' | tail -1 | tee -a "$OUT/config5_8k_n8_tolerance_sweep.jsonl"
done
