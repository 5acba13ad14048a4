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
ONCE = "--once" in sys.argv     # profiling runs: one launch per user-metric kernel on the 1080p config4 scene
# the reference's kerr_schild as a user metric in its three forms (tests/test_user_metric.py)
FORMS = [("matrix, 4 partials", "kerr_schild_as_written", ""), ("matrix, declared stationary", "kerr_schild_as_written", "#pragma rtgr stationary\n"),
         ("Kerr-Schild form (f, k), declared stationary", "kerr_schild_form", "")]
mids = []
for label, name, decl in FORMS:
    t0 = time.perf_counter()
    mids.append((label, ctx.compile_metric(decl + open(os.path.join(pkg.METRIC_SOURCES, name + ".cu")).read()), time.perf_counter() - t0))
peak = max(ctx.fp64_peak(0)[0] for _ in range(3))
for name, size in ((("config4", (1920, 1080)),) if ONCE else (("example2", None), ("config3", None), ("config4", (1920, 1080)))):
    sc = pkg.scenes.BY_NAME[name]()
    if size:
        sc = sc.with_size(*size)
    ctx.render_resident(sc)
    b = ctx.render_resident(sc) if ONCE else min((ctx.render_resident(sc) for _ in range(3)), key=lambda s: s["kernel_ms"])
    row = {"scene": sc.name, "ni": sc.ni, "nj": sc.nj, "a": sc.a, "fp64_peak_tflops": peak,
           "builtin": {"kernel_ms": b["kernel_ms"], "rays_per_s": b["rays"] / b["kernel_ms"] * 1e3, "rhs_evals": b["rhs_evals"]},
           "user_metric": {}}
    for label, mid, compile_s in mids:
        ctx.set_metric_params(mid, (sc.M, sc.a))
        scene = replace(sc, metric=mid)
        if ONCE:
            u = ctx.render_resident(scene)
        else:
            ctx.render_resident(scene)
            u = min((ctx.render_resident(scene) for _ in range(3)), key=lambda s: s["kernel_ms"])
        row["user_metric"][label] = {
            "kernel_ms": u["kernel_ms"], "rays_per_s": u["rays"] / u["kernel_ms"] * 1e3, "rhs_per_s": u["rhs_evals"] / u["kernel_ms"] * 1e3,
            "rhs_evals": u["rhs_evals"], "nvrtc_compile_s": round(compile_s, 2), "slowdown_vs_builtin": u["kernel_ms"] / b["kernel_ms"],
            # W_rhs,ref = 1781 flops: the right-hand side as the reference executes it
            "as_written_tflops": 1781 * u["rhs_evals"] / u["kernel_ms"] / 1e9}
    print(json.dumps(row), flush=True)
for _, mid, _ in mids:
    ctx.release_metric(mid)
ctx.close()
