"""Loader for csrc/libraytracegr_cuda.so.  There is no fallback: if the CUDA library is missing or
lacks a symbol the import of the product path fails loudly."""
import ctypes as C
import os
import subprocess

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = None

#: every entry point include/raytracegr_cuda.h declares
SYMBOLS = [
    "rtgr_create", "rtgr_destroy", "rtgr_last_error", "rtgr_version", "rtgr_device_count",
    "rtgr_default_params", "rtgr_alloc_pinned", "rtgr_free_pinned", "rtgr_trace_pixels", "rtgr_render",
    "rtgr_render_tiles", "rtgr_make_canvas", "rtgr_rhs_batch", "rtgr_upload_pixels", "rtgr_trace_resident",
    "rtgr_render_resident", "rtgr_fp64_peak", "rtgr_fp64_microbench",
    "rtgr_trace_canvas", "rtgr_host_register", "rtgr_host_unregister", "rtgr_host_is_pinned",
    "rtgr_trace_paths", "rtgr_metric_compile", "rtgr_metric_set_params", "rtgr_metric_release", "rtgr_metric_check",
    "rtgr_frame_create", "rtgr_frame_open", "rtgr_render_frame", "rtgr_trace_canvas_frame", "rtgr_frame_set_participants", "rtgr_frame_read", "rtgr_frame_clear", "rtgr_frame_close",
]


def library_path():
    # RTGR_LIBRARY: developer override used by the build-variant sweeps (tests/gpu_session.sh, section `variants`); it must name
    # another build of this same CUDA library -- there is no other implementation to point it at
    return os.environ.get("RTGR_LIBRARY") or os.path.join(_CSRC, "libraytracegr_cuda.so")


def kernel_source_sha16():
    """sha256 (first 16 hex digits) of the CUDA sources of the library: names the kernel build a profile under
    profiles/ belongs to (bench.py says whether the capture it quotes is of the kernel it just timed)."""
    import hashlib
    h = hashlib.sha256()
    for name in ("raytracegr_cuda.cu", "rtgr_core.cuh", "rtgr_trace.cuh", "rtgr_kernels.cuh", "rtgr_generic.cuh",
                 "rtgr_scene.h", "rtgr_jit.h", "gen_tables.py"):
        with open(os.path.join(_CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build_library():
    """Compile the CUDA library in-tree for sm_100a (make -C csrc)."""
    subprocess.check_call(["make", "-s", "-C", _CSRC])
    return library_path()


def lib():
    """dlopen the CUDA library and declare the ABI signatures."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "libraytracegr_cuda.so is not built (%s). Run `python __graft_entry__.py` or `make -C %s`. "
            "There is no CPU fallback." % (path, _CSRC))
    L = C.CDLL(path)
    missing = [s for s in SYMBOLS if not hasattr(L, s)]
    if missing:
        raise RuntimeError("libraytracegr_cuda.so lacks symbols: %s" % ", ".join(missing))
    dp, ip, u8p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    P, O, Cam, St = C.POINTER(_abi.rtgr_params), C.POINTER(_abi.rtgr_object), C.POINTER(_abi.rtgr_camera), C.POINTER(_abi.rtgr_stats)
    ctx = C.c_void_p
    L.rtgr_create.argtypes = [C.POINTER(ctx), C.POINTER(C.c_int), C.c_int]
    L.rtgr_destroy.argtypes = [ctx]
    L.rtgr_destroy.restype = None
    L.rtgr_last_error.restype = C.c_char_p
    L.rtgr_device_count.argtypes = [ctx]
    L.rtgr_default_params.argtypes = [P, C.c_int]
    L.rtgr_default_params.restype = None
    L.rtgr_alloc_pinned.argtypes = [C.c_uint64]
    L.rtgr_alloc_pinned.restype = C.c_void_p
    L.rtgr_free_pinned.argtypes = [C.c_void_p]
    L.rtgr_free_pinned.restype = None
    L.rtgr_trace_pixels.argtypes = [ctx, P, O, C.c_int, C.c_void_p, C.c_int64, dp, ip, ip, ip, St]
    L.rtgr_trace_canvas.argtypes = [ctx, P, O, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, ip, ip, ip, St]
    L.rtgr_host_register.argtypes = [C.c_void_p, C.c_uint64]
    L.rtgr_host_unregister.argtypes = [C.c_void_p]
    L.rtgr_host_is_pinned.argtypes = [C.c_void_p]
    L.rtgr_trace_paths.argtypes = [ctx, P, O, C.c_int, dp, C.c_int64, C.c_int32, dp, ip, dp, ip, ip, St]
    L.rtgr_metric_compile.argtypes = [ctx, C.c_char_p, C.POINTER(C.c_int32)]
    L.rtgr_metric_set_params.argtypes = [ctx, C.c_int32, dp, C.c_int]
    L.rtgr_metric_release.argtypes = [ctx, C.c_int32]
    L.rtgr_metric_check.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
    L.rtgr_render.argtypes = [ctx, P, O, C.c_int, Cam, u8p, dp, dp, ip, ip, ip, St]
    L.rtgr_render_tiles.argtypes = [ctx, P, O, C.c_int, Cam, C.c_int, C.c_int, u8p, dp, dp, ip, ip, ip, St]
    L.rtgr_make_canvas.argtypes = [ctx, P, Cam, C.c_void_p]
    L.rtgr_rhs_batch.argtypes = [ctx, P, dp, C.c_int64, dp]
    L.rtgr_upload_pixels.argtypes = [ctx, C.c_void_p, C.c_int64]
    L.rtgr_trace_resident.argtypes = [ctx, P, O, C.c_int, St]
    L.rtgr_render_resident.argtypes = [ctx, P, O, C.c_int, Cam, C.c_int, C.c_int, St]
    L.rtgr_frame_create.argtypes = [ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p), u8p]
    L.rtgr_frame_open.argtypes = [ctx, u8p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.rtgr_render_frame.argtypes = [C.c_void_p, P, O, C.c_int, Cam, St]
    L.rtgr_trace_canvas_frame.argtypes = [C.c_void_p, P, O, C.c_int, C.c_void_p, C.c_int, C.c_int, St]
    L.rtgr_frame_read.argtypes = [C.c_void_p, u8p]
    L.rtgr_frame_set_participants.argtypes = [C.c_void_p, C.c_int]
    L.rtgr_frame_clear.argtypes = [C.c_void_p]
    L.rtgr_frame_close.argtypes = [C.c_void_p]
    L.rtgr_frame_close.restype = None
    L.rtgr_fp64_peak.argtypes = [ctx, C.c_int, dp, dp]
    L.rtgr_fp64_microbench.argtypes = [ctx, C.c_int, C.c_int, dp, dp]
    _LIB = L
    return L


def last_error():
    return lib().rtgr_last_error().decode("utf-8", "replace")
