#!/bin/bash
# example2 at the reference's 200x200: frame time against the number of resident CTAs per SM (quick form)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/${1:-smallq}
for k in 1 2 4; do
  line=$(RTGR_CTAS_PER_SM=$k timeout 50 python bench.py --workload example2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
  echo "example2 200x200 ctas_per_sm=$k $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms %.3f ms_per_step %.3f drain_ms %.3f" % (d["kernel_ms_per_step"], d["ms_per_step"], d["roofline"]["drain_ms"]))' 2>&1)" | tee -a gpurun_out/${1:-smallq}/small_frames.log
done
