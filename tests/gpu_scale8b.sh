#!/bin/bash
# 8-GPU box: 4K bench at N=8 and N=2, config5 8K sweep at N=8 (after the shard-balancing change)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-scale8b}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for n in 8 2; do
  echo "== bench N=$n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 2>"$OUT/bench_n$n.err" | grep '^{' | tail -1 | tee "$OUT/bench_n$n.json" | cut -c1-300
done
for tol in 1e-6 1e-10; do
  echo "== config5 tol=$tol N=8"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus 8 --workload config5 --tol $tol --steps 3 --warmup 3 --no-e2e 2>"$OUT/c5_$tol.err" | grep '^{' | tail -1 | tee -a "$OUT/config5_8k_n8_tolerance_sweep.jsonl" | cut -c1-300
done
