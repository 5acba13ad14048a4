#!/bin/bash
# one full ncu capture of the trace kernel (config4 scene at 1920x1080), raw + source pages exported as CSV
set -u
cd "$(dirname "$0")/.."
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace" -f \
    python bench.py --ni 1920 --nj 1080 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/prof_cmd.log" 2>&1
ncu -i "$OUT/prof_trace.ncu-rep" --page raw --csv > "$OUT/raw.csv" 2>/dev/null
ncu -i "$OUT/prof_trace.ncu-rep" --page source --csv --print-source sass > "$OUT/source_sass.csv" 2>/dev/null
ls -la "$OUT"
