"""ctypes binding of tests/host_shim.cpp: the product's per-lane arithmetic compiled for the host
with a one-lane scheduler.  TEST SCAFFOLDING ONLY (lets the integrator logic be checked against the
oracle on machines without a GPU)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    src = os.path.join(ROOT, "tests", "host_shim.cpp")
    out = os.path.join(ROOT, "tests", "libhost_shim.so")
    csrc = os.path.join(ROOT, "raytracegr.jl_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("rtgr_core.cuh", "rtgr_trace.cuh", "rtgr_scene.h", "tsit5_tables.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-mavx2",
                           "-Wno-unknown-pragmas", "-shared", "-o", out, src])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def rhs_batch(params, states):
    states = np.ascontiguousarray(states, dtype=np.float64)
    out = np.empty_like(states)
    assert lib().shim_rhs_batch(C.byref(params), _p(states), C.c_int64(states.shape[0]), _p(out)) == 0
    return out


def make_canvas(params, cam):
    px = np.zeros((cam.ni * cam.nj, 11))
    assert lib().shim_make_canvas(C.byref(params), C.byref(cam), _p(px)) == 0
    return px


def _outs(n):
    return (np.zeros((n, 3)), np.zeros((n, 8)), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32),
            np.zeros(4, np.uint64))


def trace_pixels(params, objs, nobj, pixels):
    px = np.ascontiguousarray(pixels, dtype=np.float64)
    n = px.shape[0]
    rgb, fs, oid, st, ns, cnt = _outs(n)
    rc = lib().shim_trace_pixels(C.byref(params), objs, C.c_int(nobj), _p(px), C.c_int64(n), _p(rgb), _p(fs),
                                 _p(oid, C.c_int32), _p(st, C.c_int32), _p(ns, C.c_int32), _p(cnt, C.c_uint64))
    assert rc == 0
    return dict(rgb=rgb, final_state=fs, obj_id=oid, status=st, nsteps=ns,
                counters=dict(rays=int(cnt[0]), attempts=int(cnt[1]), accepted=int(cnt[2]), rejected=int(cnt[3])))


def render_tiles(params, objs, nobj, cam, tile_offset=0, tile_stride=1, out=None):
    n = cam.ni * cam.nj
    if out is None:
        rgb, fs, oid, st, ns, _ = _outs(n)
        out = dict(rgb8=np.zeros((cam.nj, cam.ni, 3), np.uint8), rgb=rgb, final_state=fs, obj_id=oid, status=st, nsteps=ns)
    cnt = np.zeros(4, np.uint64)
    rc = lib().shim_render_tiles(C.byref(params), objs, C.c_int(nobj), C.byref(cam), C.c_int(tile_offset),
                                 C.c_int(tile_stride), _p(out["rgb8"], C.c_uint8), _p(out["rgb"]), _p(out["final_state"]),
                                 _p(out["obj_id"], C.c_int32), _p(out["status"], C.c_int32), _p(out["nsteps"], C.c_int32),
                                 _p(cnt, C.c_uint64))
    assert rc == 0
    out["counters"] = dict(rays=int(cnt[0]), attempts=int(cnt[1]), accepted=int(cnt[2]), rejected=int(cnt[3]))
    return out


def render_frame(params, objs, nobj, cam, head_addr, rgb8):
    """The shared-queue model of rtgr_render_frame: `head_addr` is the address of a uint64 queue head other
    processes draw from too, `rgb8` the (shared) image the pixels are stored into.  Returns this caller's counters."""
    cnt = np.zeros(4, np.uint64)
    rc = lib().shim_render_frame(C.byref(params), objs, C.c_int(nobj), C.byref(cam), C.c_void_p(head_addr),
                                 _p(rgb8, C.c_uint8), _p(cnt, C.c_uint64))
    assert rc == 0
    return dict(rays=int(cnt[0]), attempts=int(cnt[1]), accepted=int(cnt[2]), rejected=int(cnt[3]))


def trace_canvas_frame(params, objs, nobj, pixels, head_addr):
    """shim_trace_canvas_frame: this caller's share of ONE (nj, ni, 11) float64 canvas that other processes trace too
    (`pixels` is a view of shared memory), rays drawn from the shared head at `head_addr`; rgb written in place."""
    nj, ni = pixels.shape[:2]
    cnt = np.zeros(4, np.uint64)
    rc = lib().shim_trace_canvas_frame(C.byref(params), objs, C.c_int(nobj), C.c_void_p(pixels.ctypes.data), C.c_int(ni), C.c_int(nj),
                                       C.c_void_p(head_addr), _p(cnt, C.c_uint64))
    assert rc == 0
    return dict(rays=int(cnt[0]), attempts=int(cnt[1]), accepted=int(cnt[2]), rejected=int(cnt[3]))
