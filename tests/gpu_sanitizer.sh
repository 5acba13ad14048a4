#!/bin/bash
# compute-sanitizer over every kernel family on small scenes (race detection / sanitizers)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-sanitizer}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_run.py > "$OUT/$tool.log" 2>&1
  tail -4 "$OUT/$tool.log"
done
