#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
fused render, in-place canvas trace (page-locked and pageable), 1-D pixels trace, ray paths, rhs batch,
make_canvas, and a run-time compiled user metric."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
ctx = pkg.Context([0])
for name in ("example1", "example2", "config4"):
    sc = pkg.scenes.BY_NAME[name](ni=70, nj=45)          # ragged: not a multiple of the 32x32 tile
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    full = ctx.render(sc, want=("rgb8", "rgb_f64", "final_state", "obj_id", "status", "nsteps"))
    px = ctx.make_canvas(p, cam)
    buf = pkg.PinnedArray((45, 70, 11))
    buf.array[...] = px.reshape(45, 70, 11)
    a = ctx.trace_canvas(p, objs, nobj, buf.array, want=("final_state", "obj_id"))
    pg = np.array(px.reshape(45, 70, 11), copy=True)
    b = ctx.trace_canvas(p, objs, nobj, pg, tile_offset=1, tile_stride=2, want=("obj_id",))
    c = ctx.trace_pixels(p, objs, nobj, np.array(px, copy=True), want=("final_state", "obj_id", "status", "nsteps"))
    assert np.array_equal(a["final_state"], full["final_state"]) and np.array_equal(c["final_state"], full["final_state"])
    r = ctx.trace_paths(p, objs, nobj, px[:64, :8], max_points=128)
    assert np.array_equal(r["final_state"], full["final_state"][:64])
    d = ctx.rhs_batch(p, px[:1000, :8])
    buf.free()
    print(name, "ok", full["stats"]["rays"], "rays", full["stats"]["rhs_evals"], "rhs evals", flush=True)
src = open(os.path.join(pkg.METRIC_SOURCES, "kerr_schild_as_written.cu")).read()
mid = ctx.compile_metric(src, par=(1.0, 0.5))
from dataclasses import replace
sc = replace(pkg.scenes.example2(ni=40, nj=33), metric=mid)
out = ctx.render(sc, want=("rgb8", "obj_id"))
p, objs, nobj, cam = pkg.scenes.to_abi(sc)
r = ctx.trace_paths(p, objs, nobj, ctx.make_canvas(p, cam)[:32, :8], max_points=64)
print("user metric ok", out["stats"]["rays"], "rays", flush=True)
ctx.release_metric(mid)
ctx.close()
print("SANITIZER_RUN_COMPLETE")
