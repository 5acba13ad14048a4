#!/bin/bash
# One GPU-box session: parity tests, smoke, bench (both arms), FP64 microbench, ncu launch list and one
# full ncu capture of the trace kernel.  Everything lands in gpurun_out/.
# usage: tests/gpu_round.sh [tag]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; grep -m1 'model name' /proc/cpuinfo >> "$OUT/host.txt"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as e; e.smoke()' 2>&1 | tail -5 | tee "$OUT/smoke.log"
echo "== fp64 microbench"; timeout 300 python tests/microbench_fp64.py 2>&1 | tail -3 | tee "$OUT/fp64_peak.log"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > "$OUT/clocks.csv" 2>&1 &
SMI=$!
echo "== bench (ours)"; timeout 900 python bench.py 2>&1 | tail -1 | tee "$OUT/bench.json"
kill $SMI
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee "$OUT/bench_reference.json"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/launches_cmd.log" 2>&1
echo "== ncu full (trace kernel, config4 scene at 1920x1080)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace" -f \
    python bench.py --ni 1920 --nj 1080 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/prof_cmd.log" 2>&1
ls -la "$OUT"
