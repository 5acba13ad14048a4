#!/bin/bash
# 8-GPU box: the default bench under torchrun at N = 8, 4, 2 (as the driver launches it)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-scale8c}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 2>"$OUT/bench_n$n.err" | grep '^{' | tail -1 > "$OUT/bench_n$n.json"
  python -c "
import json; d=json.loads(open('$OUT/bench_n$n.json').read())
print($n, 'value %.4g ms %.2f'%(d['value'],d['ms_per_step']), [round(v,1) for v in d['kernel_ms_per_rank']], 'e2e ms %.2f'%d['e2e']['ms_per_step'])"
done
