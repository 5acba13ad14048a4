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
# The kernel variants with per-warp shared-memory slots (forced here: the RGB8 patch staging, used when the image lives
# in another GPU's memory, and the chunk-wise reads / patch-wise write-back of a Pixel canvas): racecheck watches the
# warp-synchronous hand-over of the slots.  72 x 44: whole 8x4 patches in rows that are a multiple of 8 long, plus a border.
for flag in ("1", "0"):
    os.environ["RTGR_RGB8_STAGING"] = flag
    os.environ["RTGR_CHUNK_RAYS"] = flag
    for name in ("example1", "config4"):
        sc = pkg.scenes.BY_NAME[name](ni=72, nj=44)
        p, objs, nobj, cam = pkg.scenes.to_abi(sc)
        img = ctx.render(sc, want=("rgb8",))["rgb8"]
        buf = pkg.PinnedArray((44, 72, 11))
        buf.array[...] = ctx.make_canvas(p, cam).reshape(44, 72, 11)
        ctx.trace_canvas(p, objs, nobj, buf.array)
        q = np.rint(255.0 * np.clip(buf.array[:, :, 8:], 0.0, 1.0)).astype(np.uint8)
        assert np.array_equal(q, img), (flag, name)
        frame = pkg.Frame(ctx, 72, 44)
        frame.render(sc)
        assert np.array_equal(frame.read(), img), (flag, name)
        buf.array[:, :, 8:] = 0.0
        frame.trace_canvas(p, objs, nobj, buf.array)
        assert np.array_equal(np.rint(255.0 * np.clip(buf.array[:, :, 8:], 0.0, 1.0)).astype(np.uint8), img), (flag, name)
        frame.close()
        buf.free()
        print("slots", "on" if flag == "1" else "off", name, "ok", flush=True)
del os.environ["RTGR_RGB8_STAGING"], os.environ["RTGR_CHUNK_RAYS"]
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
