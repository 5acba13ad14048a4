"""GPU tests of the shared frame = the cross-GPU dynamic tile queue (rtgr_frame_*): all
participants draw 8x4-pixel patches from ONE queue head in the owner GPU's memory and store their pixels
straight into the owner's image.  The image must equal rtgr_render's bit for bit whoever traced what."""
import os
import subprocess
import sys

import numpy as np
import pytest

import conftest

pytestmark = pytest.mark.gpu
ROOT = conftest.ROOT


def _scene(pkg, name, ni, nj):
    return pkg.scenes.BY_NAME[name]().with_size(ni, nj)


@pytest.mark.parametrize("name,ni,nj", [("example2", 200, 200), ("config4", 237, 131), ("example1", 97, 64)])
def test_frame_single_participant_equals_render(pkg, ctx, name, ni, nj):
    sc = _scene(pkg, name, ni, nj)
    ref = ctx.render(sc, want=("rgb8",))
    frame = pkg.Frame(ctx, ni, nj)
    try:
        assert frame.owner and len(frame.handle) == 64
        for _ in range(3):                      # consecutive frames alternate between the two queue heads
            frame.clear()
            st = frame.render(sc)
            assert st["rays"] == ni * nj
            assert st["rhs_evals"] == ref["stats"]["rhs_evals"]
            assert np.array_equal(frame.read(), ref["rgb8"])
    finally:
        frame.close()


def test_frame_rejects_a_camera_of_another_size(pkg, ctx):
    from raytracegr_jl_b200 import host
    frame = pkg.Frame(ctx, 64, 48)
    try:
        with pytest.raises(host.RtgrError) as e:
            frame.render(_scene(pkg, "example2", 64, 64))
        assert "differ" in str(e.value)
    finally:
        frame.close()


def test_frame_two_processes_share_one_queue(pkg, ctx):
    """Owner and a second process (same GPU here; one per GPU in production) render ONE frame together
    through the IPC mapping: shares add up to the frame, the image equals the single-process render."""
    name, ni, nj, frames = "config4", 960, 540, 3
    sc = _scene(pkg, name, ni, nj)
    ref = ctx.render(sc, want=("rgb8",))
    frame = pkg.Frame(ctx, ni, nj)
    peer = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "frame_peer.py"), "0", frame.handle.hex(),
                             name, str(ni), str(nj), str(frames)],
                            stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=ROOT)
    try:
        line = peer.stdout.readline().strip()
        if line.startswith("open-failed"):
            pytest.skip("a second process cannot share this GPU / map the frame here: " + line)
        assert line == "ready", line
        shares = []
        for _ in range(frames):
            frame.clear()
            peer.stdin.write("go\n"); peer.stdin.flush()          # barrier in: both start the frame
            st = frame.render(sc)
            words = peer.stdout.readline().split()                 # barrier out: the peer's share is done
            assert words and words[0] == "done", words
            shares.append((st["rays"], int(words[1])))
            assert st["rays"] + int(words[1]) == ni * nj
            assert np.array_equal(frame.read(), ref["rgb8"])
        # with both kernels resident on one GPU the split is arbitrary; it must only be a partition
        print("shares (owner, peer):", shares)
    finally:
        try:
            peer.stdin.close()
        except Exception:
            pass
        peer.wait(timeout=60)
        frame.close()


@pytest.mark.parametrize("chunk", ["default", "1", "0"])
@pytest.mark.parametrize("name,ni,nj", [("config4", 237, 131), ("config4", 256, 96), ("example1", 97, 64)])
def test_frame_trace_canvas_single_participant(pkg, ctx, monkeypatch, name, ni, nj, chunk):
    # rtgr_trace_canvas_frame on a page-locked canvas == rtgr_trace_canvas (the trace_rays drop-in), bit for bit --
    # ray by ray, or with the rays of a patch read and its colours written back together (what a frame of four or
    # more GPUs does; forced here with RTGR_CHUNK_RAYS=1)
    if chunk != "default":
        monkeypatch.setenv("RTGR_CHUNK_RAYS", chunk)
    sc = _scene(pkg, name, ni, nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    canvas0 = ctx.make_canvas(p, cam).reshape(nj, ni, 11)
    ref = canvas0.copy()
    ctx.trace_canvas(p, objs, nobj, ref)
    frame = pkg.Frame(ctx, ni, nj)
    frame.set_participants(8 if chunk == "default" else 1)      # (the hint alone selects the patch-wise path too)
    buf = pkg.PinnedArray((nj, ni, 11))
    try:
        for _ in range(2):
            buf.array[...] = canvas0
            st = frame.trace_canvas(p, objs, nobj, buf.array)
            assert st["rays"] == ni * nj
            assert np.array_equal(buf.array, ref)
        from raytracegr_jl_b200 import host
        with pytest.raises(host.RtgrError) as e:              # pageable memory cannot be shared with the GPUs in place
            frame.trace_canvas(p, objs, nobj, canvas0.copy())
        assert "page-locked" in str(e.value)
    finally:
        buf.free()
        frame.close()


def test_frame_two_processes_share_one_host_canvas(pkg, ctx):
    """trace_rays on ONE Array{Pixel} canvas by two processes: the canvas lives in POSIX shared memory that both map
    and page-lock; the rays come from the frame's shared queue; the canvas ends up complete with no gather step."""
    name, ni, nj, frames = "config4", 480, 270, 2
    sc = _scene(pkg, name, ni, nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    canvas0 = ctx.make_canvas(p, cam).reshape(nj, ni, 11)
    ref = canvas0.copy()
    ctx.trace_canvas(p, objs, nobj, ref)
    frame = pkg.Frame(ctx, ni, nj)
    shared = pkg.SharedCanvas(nj, ni)
    peer = subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "frame_peer.py"), "0", frame.handle.hex(),
                             name, str(ni), str(nj), str(frames), shared.name],
                            stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, cwd=ROOT)
    try:
        line = peer.stdout.readline().strip()
        if line.startswith("open-failed"):
            pytest.skip("a second process cannot share this GPU / map the frame here: " + line)
        assert line == "ready", line
        for _ in range(frames):
            shared.array[...] = canvas0
            peer.stdin.write("go\n"); peer.stdin.flush()          # barrier in: both start the frame
            st = frame.trace_canvas(p, objs, nobj, shared.array)
            words = peer.stdout.readline().split()                 # barrier out: the peer's share is done
            assert words and words[0] == "done", words
            assert st["rays"] + int(words[1]) == ni * nj
            assert np.array_equal(shared.array, ref)
    finally:
        try:
            peer.stdin.close()
        except Exception:
            pass
        peer.wait(timeout=60)
        shared.close()
        frame.close()


def test_frame_multi_device_in_one_context(pkg, ctx):
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    name, ni, nj = "config4", 960, 540
    sc = _scene(pkg, name, ni, nj)
    ref = ctx.render(sc, want=("rgb8",))
    with pkg.Context(list(range(min(4, nd)))) as multi:
        frame = pkg.Frame(multi, ni, nj)
        try:
            for _ in range(2):
                frame.clear()
                st = frame.render(sc)
                assert st["rays"] == ni * nj
                assert np.array_equal(frame.read(), ref["rgb8"])
        finally:
            frame.close()


@pytest.mark.parametrize("queue", ["shared", "static"])
def test_multi_device_context_matches_single(pkg, ctx, monkeypatch, queue):
    """The devices of ONE context draw a call's tiles from one queue head in device 0's memory and write into device
    0's buffers or the caller's page-locked canvas (the default), or work on per-device tile sets whose packed tiles
    the host gathers (RTGR_MULTI_QUEUE=static) -- render, render_tiles and trace_canvas must return exactly what a
    single device returns either way.  (Needs >= 2 GPUs.)"""
    import torch
    nd = torch.cuda.device_count()
    if nd < 2:
        pytest.skip("needs >= 2 GPUs")
    sc = _scene(pkg, "config4", 480, 270)
    want = ("rgb8", "obj_id", "final_state", "nsteps")
    ref = ctx.render(sc, want=want)
    ref_part = ctx.render(sc, want=want, tile_offset=1, tile_stride=3)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    canvas0 = ctx.make_canvas(p, cam).reshape(sc.nj, sc.ni, 11)
    ref_canvas = canvas0.copy()
    ctx.trace_canvas(p, objs, nobj, ref_canvas)
    monkeypatch.setenv("RTGR_MULTI_QUEUE", queue)
    with pkg.Context(list(range(min(4, nd)))) as multi:
        out = multi.render(sc, want=want)
        for k in want:
            assert np.array_equal(out[k], ref[k]), k
        assert out["stats"]["rays"] == sc.ni * sc.nj
        part = multi.render(sc, want=want, tile_offset=1, tile_stride=3)
        for k in want:
            assert np.array_equal(part[k], ref_part[k]), k
        pageable = canvas0.copy()
        multi.trace_canvas(p, objs, nobj, pageable)
        assert np.array_equal(pageable, ref_canvas)
        buf = pkg.PinnedArray((sc.nj, sc.ni, 11))
        try:
            buf.array[...] = canvas0
            multi.trace_canvas(p, objs, nobj, buf.array)
            assert np.array_equal(buf.array, ref_canvas)
        finally:
            buf.free()
