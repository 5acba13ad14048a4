#!/bin/bash
# Build-variant sweep on the GPU box: scripts/sweep_build.sh "<NVEXTRA flags 1>" "<flags 2>" ...
# Each variant is compiled in-tree and timed with the kernel-only bench at 1920x1080 (config4 scene).
set -u
cd "$(dirname "$0")/.."
for flags in "$@"; do
  make -s -C raytracegr.jl_b200/csrc clean
  make -s -C raytracegr.jl_b200/csrc NVEXTRA="$flags" || { echo "BUILD FAILED: $flags"; continue; }
  regs=$(grep -A2 'trace_kernelILi1ELi0' raytracegr.jl_b200/csrc/ptxas.log | grep -o 'Used [0-9]* registers' | head -1)
  spill=$(grep -A1 'trace_kernelILi1ELi0' raytracegr.jl_b200/csrc/ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)
  line=$(python bench.py --ni 1920 --nj 1080 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "VARIANT [$flags] $regs, $spill :: $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("rays/s %.3e  kernel_ms %.2f  frac %.4f  acc_steps %d" % (d["value"], d["kernel_ms_per_step"], d["roofline"]["frac"], d["work"]["step_attempts"]))' 2>&1)"
done
make -s -C raytracegr.jl_b200/csrc clean; make -s -C raytracegr.jl_b200/csrc
