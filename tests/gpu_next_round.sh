#!/bin/bash
# First GPU session of the next round (1 GPU, ~12 min): everything DESIGN.md 7.2 lists for one GPU.
#   1. parity suite + smoke with the current hot loop (the three savings made after the last measurement)
#   2. default bench (both arms), ncu launch list, full ncu capture of the trace kernel  -> profiles/r02a_*
#   3. build-variant sweep: -DRTGR_ERRNORM_RCP0
#   4. small-frame sweep (RTGR_CTAS_PER_SM = 1..4 at 200x200 ... 960x540)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
bash tests/gpu_final.sh "$TAG" 2>&1 | tail -60
echo "== variant sweep (1080p)"; bash tests/sweep_build.sh "" "-DRTGR_ERRNORM_RCP0" 2>&1 | grep VARIANT | tee "$OUT/variant_sweep.log"
echo "== variant parity (RCP0)"
make -s -C raytracegr.jl_b200/csrc clean; make -s -C raytracegr.jl_b200/csrc NVEXTRA=-DRTGR_ERRNORM_RCP0
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu_rcp0.log"
bash tests/quick_bench.sh --steps 3 --warmup 3 2>&1 | tee "$OUT/quick_4k_rcp0.log"
make -s -C raytracegr.jl_b200/csrc clean; make -s -C raytracegr.jl_b200/csrc
echo "== small frames"; bash tests/gpu_small_frames.sh "$TAG" 2>&1 | tail -14
