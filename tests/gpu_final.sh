#!/bin/bash
# Round-final GPU-box session (1 GPU): parity suite, smoke, FP64 microbenchmarks, the default bench (both
# arms), the other BASELINE configurations, the ncu launch list of the bench command and one full ncu capture
# of the trace kernel on the bench workload (4K).  Everything lands in gpurun_out/<tag>/.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host.txt"; grep -m1 'model name' /proc/cpuinfo >> "$OUT/host.txt"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as e; e.smoke()' 2>&1 | tail -3 | tee "$OUT/smoke.log"
echo "== parity report"; timeout 600 python tests/parity_report.py 2>&1 | tail -4 | tee "$OUT/parity_report.jsonl"
echo "== fp64 microbench"; timeout 300 python tests/microbench_fp64.py 2>&1 | tail -12 | tee "$OUT/fp64_modes.log"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 250 > "$OUT/clocks.csv" 2>&1 &
SMI=$!
echo "== bench (ours, default = config4 4K)"; timeout 900 python bench.py 2>"$OUT/bench.err" | tail -1 | tee "$OUT/bench_config4.json"
kill $SMI
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee "$OUT/bench_reference.json"
for w in example1 example2 config3; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee "$OUT/bench_$w.json"
done
echo "== user metric bench"; timeout 600 python tests/bench_user_metric.py 2>&1 | tail -3 | tee "$OUT/bench_user_metric.jsonl"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/launches_cmd.log" 2>&1
echo "== ncu full (trace kernel, bench workload 3840x2160)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace_4k" -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/prof_cmd.log" 2>&1
ncu -i "$OUT/prof_trace_4k.ncu-rep" --page raw --csv > "$OUT/trace_kernel_4k_ncu_raw.csv" 2>/dev/null
ncu -i "$OUT/prof_trace_4k.ncu-rep" --page details --csv > "$OUT/trace_kernel_4k_ncu_details.csv" 2>/dev/null
ls -la "$OUT"
