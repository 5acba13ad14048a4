#!/bin/bash
# Second GPU-box session of a round: operand-mix microbenchmarks, the bench on every BASELINE.json
# configuration, and a full ncu capture of the trace kernel on the bench workload itself (4K).
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.log"
echo "== fp64 microbench"; timeout 300 python tests/microbench_fp64.py 2>&1 | tail -12 | tee "$OUT/fp64_modes.log"
echo "== bench config4 (default)"; timeout 900 python bench.py 2>&1 | tail -1 | tee "$OUT/bench_config4.json"
for w in example1 example2 config3; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tail -1 | tee "$OUT/bench_$w.json"
done
echo "== bench config5 (8K) tolerance sweep, kernel path only"
python - <<'PY' 2>&1 | tee "$OUT/bench_config5_sweep.log"
import json, sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as e
pkg = e.load_package()
ctx = pkg.Context([0])
for tol in (1e-6, 1e-7, 1e-8, 1e-9, 1e-10):
    sc = pkg.scenes.config5(tol=tol)
    ctx.render_resident(sc)
    best = None
    for _ in range(3):
        st = ctx.render_resident(sc)
        if best is None or st["kernel_ms"] < best["kernel_ms"]:
            best = st
    att = best["steps_accepted"] + best["steps_rejected"]
    fl = 383 * best["rhs_evals"] + 516 * att
    print(json.dumps({"workload": sc.name, "tol": tol, "rays": best["rays"], "kernel_ms": best["kernel_ms"],
                      "rays_per_s": best["rays"] / best["kernel_ms"] * 1e3, "rhs_per_s": best["rhs_evals"] / best["kernel_ms"] * 1e3,
                      "rhs_per_ray": best["rhs_evals"] / best["rays"], "rejected": best["steps_rejected"],
                      "model_tflops": fl / best["kernel_ms"] / 1e9, "drain_ms": best["drain_ms"]}))
ctx.close()
PY
echo "== ncu full (trace kernel, bench workload = config4 at 3840x2160)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 1 -c 1 -o "$OUT/prof_trace_4k" -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/prof_cmd.log" 2>&1
ls -la "$OUT"
