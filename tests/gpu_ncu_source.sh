#!/bin/bash
# Developer tool (GPU box): one full ncu capture of the trace kernel with the per-instruction source page.
# usage: tests/gpu_ncu_source.sh <tag> [bench args, default 1920x1080 of the config4 scene]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-ncu}; shift || true
ARGS=${*:---ni 1920 --nj 1080}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace" -f \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/prof_cmd.log" 2>&1
ncu -i "$OUT/prof_trace.ncu-rep" --page raw --csv > "$OUT/raw.csv" 2>/dev/null
ncu -i "$OUT/prof_trace.ncu-rep" --page source --csv --print-source sass > "$OUT/source_sass.csv" 2>/dev/null
rm -f "$OUT/prof_trace.ncu-rep"
ls -la "$OUT"
