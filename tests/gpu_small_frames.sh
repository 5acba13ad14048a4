#!/bin/bash
# Small frames (fewer rays than the GPU has threads): how does the frame time depend on the number of
# resident CTAs per SM?  tests/gpu_small_frames.sh [tag]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-small}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for spec in "example2 200 200" "config4 480 270" "config4 960 540"; do
  set -- $spec
  for k in 1 2 3 4; do
    line=$(RTGR_CTAS_PER_SM=$k timeout 300 python bench.py --workload $1 --ni $2 --nj $3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
    echo "$1 $2x$3 ctas_per_sm=$k $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms %.3f ms_per_step %.3f rays/s %.4e drain_ms %.3f" % (d["kernel_ms_per_step"], d["ms_per_step"], d["value"], d["roofline"]["drain_ms"]))' 2>&1)" | tee -a "$OUT/small_frames.log"
  done
done
