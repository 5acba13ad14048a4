"""ctypes binding of the CPU oracle (oracle/librtgr_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "librtgr_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def metric(params, x):
    g = np.zeros(16)
    lib().oracle_metric(C.byref(params), _p(np.ascontiguousarray(x, dtype=np.float64)), _p(g))
    return g.reshape(4, 4)


def dmetric(params, x):
    g = np.zeros(16); dg = np.zeros(64)
    lib().oracle_dmetric(C.byref(params), _p(np.ascontiguousarray(x, dtype=np.float64)), _p(g), _p(dg))
    return g.reshape(4, 4), dg.reshape(4, 4, 4)


def christoffel(params, x):
    G = np.zeros(64)
    lib().oracle_christoffel(C.byref(params), _p(np.ascontiguousarray(x, dtype=np.float64)), _p(G))
    return G.reshape(4, 4, 4)


def inverse4(A):
    B = np.zeros(16); det = C.c_double()
    lib().oracle_inverse4(_p(np.ascontiguousarray(A, dtype=np.float64)), _p(B), C.byref(det))
    return B.reshape(4, 4), det.value


def ks_checks_f32(params, x):
    out = np.zeros(5, dtype=np.float32)
    xs = np.ascontiguousarray(x, dtype=np.float32)
    lib().oracle_ks_checks_f32(C.byref(params), _p(xs, C.c_float), _p(out, C.c_float))
    return out


def rhs_batch(params, states, extended=False):
    states = np.ascontiguousarray(states, dtype=np.float64)
    out = np.empty_like(states)
    fn = lib().oracle_rhs_batch_ld if extended else lib().oracle_rhs_batch
    fn(C.byref(params), _p(states), C.c_int64(states.shape[0]), _p(out))
    return out


def make_canvas(params, cam):
    px = np.zeros((cam.ni * cam.nj, 11))
    lib().oracle_make_canvas(C.byref(params), C.byref(cam), _p(px))
    return px


def trace_pixels(params, objs, nobj, pixels, nthreads=0, extended=False):
    """Returns dict(pixels, final_state, obj_id, status, nsteps, stats)."""
    from importlib import import_module
    px = np.array(pixels, dtype=np.float64, order="C", copy=True)
    n = px.shape[0]
    fs = np.zeros((n, 8)); oid = np.zeros(n, dtype=np.int32); st = np.zeros(n, dtype=np.int32)
    ns = np.zeros(n, dtype=np.int32)
    import sys
    abi = sys.modules["raytracegr_jl_b200"]._abi
    stats = abi.rtgr_stats()
    fn = lib().oracle_trace_pixels_ld if extended else lib().oracle_trace_pixels
    rc = fn(C.byref(params), objs, C.c_int(nobj), _p(px), C.c_int64(n), _p(fs), _p(oid, C.c_int32),
            _p(st, C.c_int32), _p(ns, C.c_int32), C.byref(stats), C.c_int(nthreads))
    assert rc == 0
    return dict(pixels=px, final_state=fs, obj_id=oid, status=st, nsteps=ns, stats=stats.as_dict())


def quantize(pixels, ni, nj):
    out = np.zeros((nj, ni, 3), dtype=np.uint8)
    lib().oracle_quantize(_p(np.ascontiguousarray(pixels)), C.c_int(ni), C.c_int(nj), _p(out, C.c_uint8))
    return out


def num_threads():
    return int(lib().oracle_num_threads())
