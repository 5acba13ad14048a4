#!/bin/bash
# The one GPU-box session script.  usage: tests/gpu_session.sh <tag> <section> [<section> ...]
# Everything lands in gpurun_out/<tag>/.  Sections (1 GPU unless said otherwise):
#   info       nvidia-smi / host facts
#   tests      pytest -m gpu
#   smoke      __graft_entry__.smoke()
#   parity     tests/parity_report.py (ids / states / RGB of the BASELINE configurations against the oracle)
#   micro      FP64 operand-mix and latency microbenchmarks
#   bench      bench.py, default workload (config4, 4K), with the nvidia-smi clock log beside it
#   ref        bench.py --impl reference
#   workloads  bench.py on example1 / example2 / config3 (kernel + e2e, no CPU leg)
#   quick      kernel-only bench of the 4K frame, one summary line (QUICK_ARGS overrides the bench arguments)
#   variants   kernel-only bench of every prebuilt build_variants/*.so (tests/build_variant.sh)
#   user       tests/bench_user_metric.py (run-time compiled metrics against the built-in kernel)
#   small      small frames against the number of resident CTAs per SM
#   launches   ncu launch list (gpu__time_duration) of the bench command
#   ncu        one `ncu --set full` capture of trace_kernel on the bench workload: raw / details / source CSV pages
#   ncu_user   the same for the run-time compiled user-metric kernel (config4 scene at 1080p)
#   sanitizer  compute-sanitizer memcheck / racecheck / initcheck / synccheck on small scenes
#   peer_stores  2-GPU box: store requests / NVLink bytes of a frame rendered entirely into ANOTHER GPU's memory, with and
#              without the RGB8 patch staging (tests/peer_store_probe.py)
#   multi      N-GPU box (N = all visible GPUs): the >= 2-GPU tests, bench.py under torchrun at N = 2 .. all (both
#              arms at the largest N), one-process multi-device render
set -u
cd "$(dirname "$0")/.."
TAG=${1:?tag}; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
NGPU=$(nvidia-smi -L | wc -l)

summary='import sys,json
d=json.loads(sys.stdin.read())
print("kernel_ms %.2f ms_per_step %.2f rays/s %.4e drain %.2f model_frac %.4f attempts %d clocks %s" % (d["kernel_ms_per_step"], d["ms_per_step"], d["value"], d["roofline"]["drain_ms"], d["roofline"]["frac"], d["work"]["step_attempts"], d["clocks"]))'

torchrun_bench() {   # <N> <port> <bench args...>
  local n=$1 port=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port "$port" bench.py --gpus "$n" "$@"
}

for S in "$@"; do
  echo "===== $S"
  case $S in
  info)
    nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1
    nvidia-smi topo -m 2>/dev/null | head -14 >> "$OUT/gpu.txt"
    nproc > "$OUT/host.txt"; grep -m1 'model name' /proc/cpuinfo >> "$OUT/host.txt"; cat "$OUT/gpu.txt" ;;
  tests)
    timeout 1800 python -m pytest tests -m gpu -x -q -rs 2>&1 | tail -12 | tee "$OUT/pytest_gpu.log" ;;
  smoke)
    timeout 300 python -c 'import __graft_entry__ as e; e.smoke()' 2>&1 | tail -3 | tee "$OUT/smoke.log" ;;
  parity)
    timeout 900 python tests/parity_report.py 2>&1 | tail -6 | tee "$OUT/parity_report.jsonl" ;;
  micro)
    timeout 300 python tests/microbench_fp64.py 2>&1 | tail -16 | tee "$OUT/fp64_modes.log"
    timeout 300 python tests/microbench_fp64_latency.py 2>&1 | tail -8 | tee "$OUT/fp64_latency.log" ;;
  bench)
    nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 250 > "$OUT/clocks.csv" 2>&1 &
    SMI=$!
    timeout 900 python bench.py 2>"$OUT/bench.err" | tail -1 | tee "$OUT/bench_config4.json"
    kill $SMI; tail -3 "$OUT/bench.err" ;;
  ref)
    timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee "$OUT/bench_reference.json" ;;
  workloads)
    for w in example1 example2 config3; do
      timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee "$OUT/bench_$w.json"
    done ;;
  quick)
    timeout 600 python bench.py --no-e2e --no-cpu-baseline ${QUICK_ARGS:---steps 3 --warmup 3} 2>&1 | tail -1 | python -c "$summary" 2>&1 | tee -a "$OUT/quick.log" ;;
  variants)
    for so in build_variants/*.so; do
      line=$(RTGR_LIBRARY=$PWD/$so timeout 600 python bench.py --no-e2e --no-cpu-baseline ${QUICK_ARGS:---steps 3 --warmup 3} 2>&1 | tail -1)
      echo "VARIANT $(basename "$so" .so) :: $(echo "$line" | python -c "$summary" 2>&1)" | tee -a "$OUT/variants.log"
    done ;;
  user)
    timeout 900 python tests/bench_user_metric.py 2>&1 | tail -4 | tee "$OUT/bench_user_metric.jsonl" ;;
  small)
    for spec in "example2 200 200" "config4 480 270" "config4 960 540"; do
      set -- $spec
      for k in 1 2 4; do
        line=$(RTGR_CTAS_PER_SM=$k timeout 300 python bench.py --workload $1 --ni $2 --nj $3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1)
        echo "$1 $2x$3 ctas_per_sm=$k $(echo "$line" | python -c "$summary" 2>&1)" | tee -a "$OUT/small_frames.log"
      done
    done ;;
  launches)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$OUT/launches.csv" \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > "$OUT/launches_cmd.log" 2>&1
    grep -c trace_kernel "$OUT/launches.csv" ;;
  ncu)
    # the bench line of the same build right before the capture (tools/ncu_summary.py pairs the two)
    timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 > "$OUT/bench_for_ncu.json"
    timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace_4k" -f \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-l2-flush > "$OUT/prof_cmd.log" 2>&1
    ncu -i "$OUT/prof_trace_4k.ncu-rep" --page raw --csv > "$OUT/trace_kernel_4k_ncu_raw.csv" 2>/dev/null
    ncu -i "$OUT/prof_trace_4k.ncu-rep" --page details --csv > "$OUT/trace_kernel_4k_ncu_details.csv" 2>/dev/null
    ncu -i "$OUT/prof_trace_4k.ncu-rep" --page source --csv --print-source sass > "$OUT/trace_kernel_4k_source_sass.csv" 2>/dev/null
    rm -f "$OUT/prof_trace_4k.ncu-rep"; ls -la "$OUT" ;;
  ncu_user)
    timeout 1500 ncu --set full --clock-control none -k regex:rtgr_user_trace -c 3 -o "$OUT/prof_user" -f \
        python tests/bench_user_metric.py --once > "$OUT/prof_user_cmd.log" 2>&1
    ncu -i "$OUT/prof_user.ncu-rep" --page raw --csv > "$OUT/user_trace_ncu_raw.csv" 2>/dev/null
    ncu -i "$OUT/prof_user.ncu-rep" --page details --csv > "$OUT/user_trace_ncu_details.csv" 2>/dev/null
    rm -f "$OUT/prof_user.ncu-rep"; tail -3 "$OUT/prof_user_cmd.log" ;;
  sanitizer)
    for tool in memcheck racecheck initcheck synccheck; do
      timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_run.py > "$OUT/sanitizer_$tool.log" 2>&1
      tail -3 "$OUT/sanitizer_$tool.log"
    done ;;
  multi)
    nvidia-smi -L | tee "$OUT/gpus.txt"
    timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_parity.py -m gpu -q -rs -k "frame or multi_device" 2>&1 | tail -8 | tee "$OUT/pytest_multi.log"
    n=${MULTI_MIN_N:-2}
    while [ $n -le "$NGPU" ]; do
      timeout 900 bash -c "$(declare -f torchrun_bench); torchrun_bench $n 29517 --steps 8 --warmup 3" 2>"$OUT/bench_n$n.err" | grep '^{' | tail -1 | tee "$OUT/bench_n$n.json"
      tail -2 "$OUT/bench_n$n.err"
      n=$((n*2))
    done
    timeout 600 bash -c "$(declare -f torchrun_bench); torchrun_bench $NGPU 29518 --impl reference --steps 1 --warmup 0" 2>&1 | grep '^{' | tail -1 | tee "$OUT/bench_reference_n$NGPU.json"
    timeout 600 python tests/multi_device_render.py 2>&1 | tail -6 | tee "$OUT/multi_device_render.log"
    # the shared frame at all GPUs with and without the RGB8 patch staging of the remote ranks (kernel path only)
    for flag in 1 0; do
      RTGR_RGB8_STAGING=$flag timeout 600 bash -c "$(declare -f torchrun_bench); torchrun_bench $NGPU 29519 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline" 2>/dev/null | grep '^{' | tail -1 \
        | python -c "$summary" 2>&1 | sed "s/^/RTGR_RGB8_STAGING=$flag N=$NGPU :: /" | tee -a "$OUT/staging_n$NGPU.log"
    done ;;
  peer_stores)
    timeout 900 python tests/peer_store_probe.py "$OUT" 2>&1 | tail -4 | tee "$OUT/peer_store_probe.jsonl" ;;
  *) echo "unknown section $S" ;;
  esac
done
