#!/bin/bash
# usage: tests/quick_bench.sh [bench args]; prints a one-line summary of the kernel-only bench
python bench.py --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c '
import sys,json
d=json.loads(sys.stdin.read())
print("rays/s %.4e kernel_ms %s drain %.2f frac %.4f attempts %d clocks %s" % (d["value"], ["%.1f"%v for v in d["kernel_ms_per_rank"]], d["roofline"]["drain_ms"], d["roofline"]["frac"], d["work"]["step_attempts"], d["clocks"]))'
