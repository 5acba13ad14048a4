#!/usr/bin/env python3
"""ONE process, all GPUs of the box in one context (what a Julia `context()` gets): the 4K frame through
rtgr_render (RGB8 image to the host) and rtgr_trace_canvas (page-locked host canvas, in place) with the shared
queue (default) and the static deal, against one device; sha256 of every image.  Developer measurement
(tests/gpu_session.sh multi); one JSON line per row."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import torch  # noqa: E402

pkg = entry.load_package()
nd = torch.cuda.device_count()
sc = pkg.scenes.BY_NAME[sys.argv[1] if len(sys.argv) > 1 else "config4"]()
p, objs, nobj, cam = pkg.scenes.to_abi(sc)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def timed(fn, reps=4):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return 1e3 * best


for devs, queue in [([0], "shared")] + [(list(range(n)), q) for n in sorted({2, nd} - {1}) if n <= nd for q in ("shared", "static")]:
    os.environ["RTGR_MULTI_QUEUE"] = queue
    with pkg.Context(devs) as ctx:
        out = {"rgb8": np.zeros((sc.nj, sc.ni, 3), dtype=np.uint8)}
        ms_render = timed(lambda: ctx.render(sc, want=("rgb8",), out=out))
        st = out["stats"]
        buf = pkg.PinnedArray((sc.nj, sc.ni, 11))
        canvas0 = ctx.make_canvas(p, cam).reshape(sc.nj, sc.ni, 11)
        buf.array[...] = canvas0

        def canvas_call():
            buf.array[:, :, 8:] = 0.0
            return ctx.trace_canvas(p, objs, nobj, buf.array)
        ms_canvas = timed(canvas_call, reps=2)   # (includes the host-side zeroing of the rgb fields: ~60 ms at 4K)
        t0 = time.perf_counter(); buf.array[:, :, 8:] = 0.0; zero_ms = 1e3 * (time.perf_counter() - t0)
        ctx.trace_canvas(p, objs, nobj, buf.array)
        q8 = np.rint(255.0 * np.clip(buf.array[:, :, 8:], 0.0, 1.0)).astype(np.uint8)
        print(json.dumps({"devices": len(devs), "queue": queue if len(devs) > 1 else "one device",
                          "render_rgb8_ms": round(ms_render, 2), "kernel_ms": round(st["kernel_ms"], 2), "rgb8_sha": sha(out["rgb8"]),
                          "trace_canvas_ms": round(ms_canvas - zero_ms, 2), "canvas_rgb8_sha": sha(q8)}), flush=True)
        buf.free()
