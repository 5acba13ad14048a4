#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/$1
for size in "" "--ni 1920 --nj 1080" "--ni 960 --nj 540"; do
  for m in row impact; do
    export RTGR_TILE_ORDER=$m
    echo "[$size] order=$m :: $(bash tests/quick_bench.sh $size --steps 3 --warmup 2 2>&1 | tail -1 | cut -c1-110)"
  done
done 2>&1 | tee gpurun_out/$1/tile_order.log
