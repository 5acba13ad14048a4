#!/bin/bash
# Developer tool (GPU box): time every prebuilt build_variants/*.so with the kernel-only bench.
# usage: tests/variant_bench.sh [bench args, default: the 4K frame, 3 steps]
set -u
cd "$(dirname "$0")/.."
ARGS=${*:---steps 3 --warmup 3}
for so in build_variants/*.so; do
  line=$(RTGR_LIBRARY=$PWD/$so python bench.py --no-e2e --no-cpu-baseline $ARGS 2>&1 | tail -1)
  echo "VARIANT $(basename $so .so) :: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("kernel_ms %.2f  rays/s %.4e  attempts %d  rhs %d" % (d["kernel_ms_per_step"], d["value"], d["work"]["step_attempts"], d["work"]["rhs_evals"]))' 2>&1)"
done
