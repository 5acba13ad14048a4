#!/usr/bin/env python3
"""Throughput of the user-metric (generic, run-time compiled) path next to the hand-specialised
built-in kernel on the same scenes.  One JSON line per scene.  W_rhs,ref = 1781 flops is the RHS as the
reference executes it; the generic path evaluates exactly that form."""
import json
import os
import sys
import time
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
ctx = pkg.Context([0])
src = open(os.path.join(pkg.METRIC_SOURCES, "kerr_schild_as_written.cu")).read()
t0 = time.perf_counter()
mid = ctx.compile_metric(src)
compile_s = time.perf_counter() - t0
peak = max(ctx.fp64_peak(0)[0] for _ in range(3))
ONCE = "--once" in sys.argv     # profiling runs: one launch of the user-metric kernel on the 1080p config4 scene
for name, size in ((("config4", (1920, 1080)),) if ONCE else (("example2", None), ("config3", None), ("config4", (1920, 1080)))):
    sc = pkg.scenes.BY_NAME[name]()
    if size:
        sc = sc.with_size(*size)
    ctx.set_metric_params(mid, (sc.M, sc.a))
    res = {}
    for label, scene in (("builtin", sc), ("user_metric", replace(sc, metric=mid))):
        if ONCE and label == "user_metric":
            res[label] = ctx.render_resident(scene)
            continue
        ctx.render_resident(scene)
        best = min((ctx.render_resident(scene) for _ in range(3)), key=lambda s: s["kernel_ms"])
        res[label] = best
    u, b = res["user_metric"], res["builtin"]
    print(json.dumps({
        "scene": sc.name, "ni": sc.ni, "nj": sc.nj, "a": sc.a, "nvrtc_compile_s": round(compile_s, 2),
        "user_metric": {"kernel_ms": u["kernel_ms"], "rays_per_s": u["rays"] / u["kernel_ms"] * 1e3,
                        "rhs_per_s": u["rhs_evals"] / u["kernel_ms"] * 1e3, "rhs_evals": u["rhs_evals"],
                        "as_written_tflops": 1781 * u["rhs_evals"] / u["kernel_ms"] / 1e9,
                        "frac_of_fp64_peak_as_written_flops": 1781 * u["rhs_evals"] / u["kernel_ms"] / 1e9 / peak},
        "builtin": {"kernel_ms": b["kernel_ms"], "rays_per_s": b["rays"] / b["kernel_ms"] * 1e3, "rhs_evals": b["rhs_evals"]},
        "slowdown_vs_builtin": u["kernel_ms"] / b["kernel_ms"], "fp64_peak_tflops": peak}), flush=True)
ctx.release_metric(mid)
ctx.close()
