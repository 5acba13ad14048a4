#!/usr/bin/env python3
"""A second PROCESS taking part in a shared frame (tests/test_gpu_frame.py; also usable by hand).

    python tests/frame_peer.py <device> <handle hex> <workload> <ni> <nj> <frames> [<shared canvas name>]

Opens the owner's frame through its IPC handle, then for every frame waits for a line on stdin (the
caller's barrier), renders its share and prints `done <rays> <kernel_ms>`.  With a shared-canvas name it
attaches to that POSIX shared-memory Pixel canvas and takes part in rtgr_trace_canvas_frame instead.
Test infrastructure."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def main():
    dev, handle, workload, ni, nj, frames = int(sys.argv[1]), bytes.fromhex(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    pkg = entry.load_package()
    scene = pkg.scenes.BY_NAME[workload]().with_size(ni, nj)
    ctx = None
    try:
        ctx = pkg.Context([dev])          # fails e.g. on a GPU in exclusive-process compute mode
        frame = pkg.Frame(ctx, ni, nj, handle=handle)
    except Exception as e:   # reported to the parent, which decides between skip and failure
        print("open-failed %s" % str(e).replace("\n", " "), flush=True)
        if ctx is not None:
            ctx.close()
        return
    canvas = pkg.SharedCanvas(nj, ni, name=sys.argv[7]) if len(sys.argv) > 7 else None
    p, objs, nobj, _cam = pkg.scenes.to_abi(scene)
    print("ready", flush=True)
    for _ in range(frames):
        if not sys.stdin.readline():
            break
        st = frame.trace_canvas(p, objs, nobj, canvas.array) if canvas is not None else frame.render(scene)
        print("done %d %.3f" % (st["rays"], st["kernel_ms"]), flush=True)
    if canvas is not None:
        canvas.close()
    frame.close()
    ctx.close()


if __name__ == "__main__":
    main()
