#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/${1:-smallq}
for spec in "480 270 1" "480 270 4" "960 540 1" "960 540 2" "960 540 4"; do
  set -- $spec
  line=$(RTGR_CTAS_PER_SM=$3 timeout 30 python bench.py --workload config4 --ni $1 --nj $2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "config4 $1x$2 ctas_per_sm=$3 $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms %.3f ms_per_step %.3f drain_ms %.3f" % (d["kernel_ms_per_step"], d["ms_per_step"], d["roofline"]["drain_ms"]))' 2>&1)" | tee -a gpurun_out/r01zz_small/small_frames.log
done
