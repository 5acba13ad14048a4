"""Host-side mirror of the reference's scene API on top of the C ABI.

Names, argument meaning and behaviour follow src/RayTraceGR.jl: `minkowski` / `kerr_schild` select
the metric (src:262, :274), `Sphere` / `Plane` are the objects (src:394-413), `make_canvas`
(src:458-478) builds the screen, `trace_rays(metric, objs, canvas)` (src:483-536) is the hot path
and returns a NEW canvas leaving its input untouched, `example1()` / `example2()` (src:542-612)
write `scenes/sphere.png` / `scenes/sphere2.png`.  All computation happens in
libraytracegr_cuda on the GPU; nothing here falls back to the CPU.
"""
import ctypes as C
import os
import struct
import zlib
from dataclasses import dataclass

import numpy as np

from . import _abi, scenes
from ._lib import last_error, lib


class RtgrError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise RtgrError(last_error())


# ---- metrics: the reference passes Julia functions; here they are tags with optional M, a -------
@dataclass(frozen=True)
class Metric:
    kind: int
    M: float = 1.0
    a: float = 0.0
    r_formula: int = _abi.RTGR_R_AS_WRITTEN

    def __call__(self, M=None, a=None, r_formula=None):
        return Metric(self.kind, self.M if M is None else M, self.a if a is None else a,
                      self.r_formula if r_formula is None else r_formula)


minkowski = Metric(_abi.RTGR_MINKOWSKI)
#: reference values M = 1, a = 0 (src:275-276); kerr_schild(a=0.9) gives a spinning hole
kerr_schild = Metric(_abi.RTGR_KERR_SCHILD)


METRIC_SOURCES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics")


def check_metric_source(source):
    """rtgr_metric_check: compile a user metric for sm_100a without touching a device; returns the
    compiler log, raises RtgrError with the diagnostics if it does not compile."""
    log = C.create_string_buffer(1 << 16)
    rc = lib().rtgr_metric_check(source.encode() if isinstance(source, str) else source, log, len(log))
    if rc != 0:
        raise RtgrError(last_error())
    return log.value.decode("utf-8", "replace")


def user_metric(source, par=(), ctx=None):
    """A Metric backed by user-supplied CUDA C++ source (see include/raytracegr_cuda.h, "user-supplied
    metrics"); usable wherever `minkowski` / `kerr_schild` are: make_canvas, trace_rays, ..."""
    ctx = ctx or default_context()
    return Metric(ctx.compile_metric(source, par))


# ---- objects -------------------------------------------------------------------------------------
@dataclass
class Plane:      # src:394-397
    time: float


@dataclass
class Sphere:     # src:409-413
    pos: tuple
    vel: tuple
    radius: float


def _marshal_objects(objs):
    if len(objs) > _abi.RTGR_MAX_OBJECTS:
        raise RtgrError("at most %d objects" % _abi.RTGR_MAX_OBJECTS)
    arr = (_abi.rtgr_object * max(1, len(objs)))()
    for i, o in enumerate(objs):
        if isinstance(o, Plane):
            arr[i].kind = _abi.RTGR_PLANE
            arr[i].time = float(o.time)
        elif isinstance(o, Sphere):
            arr[i].kind = _abi.RTGR_SPHERE
            arr[i].pos[:] = [float(v) for v in o.pos]
            arr[i].vel[:] = [float(v) for v in o.vel]
            arr[i].radius = float(o.radius)
        else:
            raise RtgrError("Called distance on abstract object")   # src:384-386
    return arr


# ---- canvas --------------------------------------------------------------------------------------
class PinnedArray:
    """A float64 numpy array in page-locked host memory (rtgr_alloc_pinned).  The GPU reads and writes
    such a buffer in place (rtgr_trace_canvas, zero copy).  Keep this object alive while `array` is used."""

    def __init__(self, shape):
        n = int(np.prod(shape))
        self._raw = lib().rtgr_alloc_pinned(max(8, 8 * n))
        if not self._raw:
            raise RtgrError(last_error())
        self.array = np.ctypeslib.as_array((C.c_double * n).from_address(self._raw)).reshape(shape)

    def free(self):
        if self._raw:
            self.array = None
            lib().rtgr_free_pinned(self._raw)
            self._raw = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SharedCanvas:
    """One Pixel canvas in POSIX shared memory, mapped and page-locked (rtgr_host_register) by every process that
    opens it: the host array of Frame.trace_canvas when the participants of a frame are separate processes (one
    per GPU).  `SharedCanvas(nj, ni)` creates it (`.name` goes to the other processes), `SharedCanvas(nj, ni,
    name=...)` attaches.  `.array` is the (nj, ni, 11) float64 view; close() unmaps (the creator also unlinks)."""

    def __init__(self, nj, ni, name=None):
        from multiprocessing import shared_memory
        nbytes = int(nj) * int(ni) * 88
        self.creator = name is None
        self._shm = shared_memory.SharedMemory(create=True, size=nbytes) if self.creator else shared_memory.SharedMemory(name=name)
        if self.creator:
            # reserve the pages now: a /dev/shm too small for the canvas (container default: 64 MB) must fail here
            # with an exception, not later with SIGBUS on first touch
            try:
                os.posix_fallocate(self._shm._fd, 0, nbytes)
            except OSError as e:
                self._shm.close()
                self._shm.unlink()
                self._shm = None
                raise RtgrError("SharedCanvas: cannot reserve %d bytes of POSIX shared memory (%s)" % (nbytes, e))
        if not self.creator:    # attaching must not hand the segment's lifetime to this process's resource tracker
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.name = self._shm.name
        self.array = np.ndarray((int(nj), int(ni), 11), dtype=np.float64, buffer=self._shm.buf)
        self._addr = self.array.ctypes.data
        self._nbytes = nbytes
        if lib().rtgr_host_register(self._addr, nbytes) != 0:
            err = RtgrError(last_error())
            self.array = None
            self._shm.close()
            if self.creator:
                self._shm.unlink()
            self._shm = None
            raise err

    def close(self):
        if self._shm is not None:
            lib().rtgr_host_unregister(self._addr)
            self.array = None
            self._shm.close()
            if self.creator:
                self._shm.unlink()
            self._shm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Canvas:
    """Canvas{Float64}: `pixels` is an (nj, ni, 11) float64 array whose memory is the reference's
    column-major Array{Pixel{Float64},2} (pixels[j, i] <-> Julia pixels[i+1, j+1]); the last axis is
    pos[4], normal[4], rgb[3] (src:446-455)."""

    def __init__(self, pixels, owner=None):
        pixels = np.ascontiguousarray(pixels, dtype=np.float64)
        assert pixels.ndim == 3 and pixels.shape[2] == 11
        self.pixels = pixels
        self._owner = owner   # PinnedArray that backs `pixels`, if any

    @property
    def ni(self):
        return self.pixels.shape[1]

    @property
    def nj(self):
        return self.pixels.shape[0]

    @property
    def rgb(self):
        return self.pixels[:, :, 8:11]

    def image8(self):
        """The 8-bit image the reference's example functions save: row = j, col = i (the transposes at
        src:566-569), value = round(255 x)."""
        return np.rint(255.0 * np.clip(self.rgb, 0.0, 1.0)).astype(np.uint8)


class Context:
    """Owns an rtgr_ctx (device buffers, streams) on the given CUDA devices."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            rc = lib().rtgr_create(C.byref(self._h), None, 0)
        else:
            ids = (C.c_int * len(devices))(*devices)
            rc = lib().rtgr_create(C.byref(self._h), ids, len(devices))
        _check(rc)

    def close(self):
        if self._h:
            for fr in list(getattr(self, "_frames", ())):   # rtgr_destroy closes them: drop the dead handles
                fr._h = C.c_void_p()
            self._frames = []
            lib().rtgr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_devices(self):
        return lib().rtgr_device_count(self._h)

    # -- hot path -------------------------------------------------------------------------------
    def trace_pixels(self, params, objs_arr, n_objs, pixels, want=()):
        """rtgr_trace_pixels on an (n, 11) float64 array (modified in place: rgb written)."""
        n = pixels.shape[0]
        assert pixels.flags.c_contiguous and pixels.dtype == np.float64
        out = {}
        fs = np.empty((n, 8)) if "final_state" in want else None
        oid = np.empty(n, dtype=np.int32) if "obj_id" in want else None
        st = np.empty(n, dtype=np.int32) if "status" in want else None
        ns = np.empty(n, dtype=np.int32) if "nsteps" in want else None
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_trace_pixels(self._h, C.byref(params), objs_arr, n_objs, pixels.ctypes.data, n,
                                       _dp(fs), _ip(oid), _ip(st), _ip(ns), C.byref(stats)))
        out.update(final_state=fs, obj_id=oid, status=st, nsteps=ns, stats=stats.as_dict())
        return out

    def trace_canvas(self, params, objs_arr, n_objs, pixels, tile_offset=0, tile_stride=1, want=()):
        """rtgr_trace_canvas on an (nj, ni, 11) float64 array, modified in place (rgb written); zero
        copy when the array lives in page-locked memory (PinnedArray / rtgr_host_register)."""
        assert pixels.ndim == 3 and pixels.shape[2] == 11 and pixels.flags.c_contiguous and pixels.dtype == np.float64
        nj, ni = pixels.shape[:2]
        n = ni * nj
        fs = np.zeros((n, 8)) if "final_state" in want else None
        oid = np.zeros(n, dtype=np.int32) if "obj_id" in want else None
        st = np.zeros(n, dtype=np.int32) if "status" in want else None
        ns = np.zeros(n, dtype=np.int32) if "nsteps" in want else None
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_trace_canvas(self._h, C.byref(params), objs_arr, n_objs, pixels.ctypes.data, ni, nj,
                                       tile_offset, tile_stride, _dp(fs), _ip(oid), _ip(st), _ip(ns), C.byref(stats)))
        return dict(final_state=fs, obj_id=oid, status=st, nsteps=ns, stats=stats.as_dict())

    def trace_paths(self, params, objs_arr, n_objs, states0, max_points=4096):
        """rtgr_trace_paths: every accepted step of n rays; returns dict(paths (n, max_points, 9), npoints,
        final_state, obj_id, status, stats)."""
        states0 = np.ascontiguousarray(states0, dtype=np.float64).reshape(-1, 8)
        n = states0.shape[0]
        paths = np.zeros((n, max_points, 9))
        npts = np.zeros(n, dtype=np.int32)
        fs, oid, st = np.zeros((n, 8)), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_trace_paths(self._h, C.byref(params), objs_arr, n_objs, _dp(states0), n, max_points,
                                      _dp(paths), _ip(npts), _dp(fs), _ip(oid), _ip(st), C.byref(stats)))
        return dict(paths=paths, npoints=npts, final_state=fs, obj_id=oid, status=st, stats=stats.as_dict())

    def render(self, scene, want=("rgb8",), tile_offset=0, tile_stride=1, out=None):
        """rtgr_render_tiles for a scenes.Scene; returns dict of requested arrays + stats."""
        p, objs, nobj, cam = scenes.to_abi(scene)
        n = scene.ni * scene.nj
        res = out if out is not None else {}
        def buf(name, shape, dtype):
            if name not in want:
                return None
            if name not in res or res[name] is None:
                res[name] = np.zeros(shape, dtype=dtype)
            return res[name]
        rgb8 = buf("rgb8", (scene.nj, scene.ni, 3), np.uint8)
        rgbf = buf("rgb_f64", (n, 3), np.float64)
        fs = buf("final_state", (n, 8), np.float64)
        oid = buf("obj_id", (n,), np.int32)
        st = buf("status", (n,), np.int32)
        ns = buf("nsteps", (n,), np.int32)
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_render_tiles(self._h, C.byref(p), objs, nobj, C.byref(cam), tile_offset, tile_stride,
                                       _u8p(rgb8), _dp(rgbf), _dp(fs), _ip(oid), _ip(st), _ip(ns), C.byref(stats)))
        res["stats"] = stats.as_dict()
        return res

    def render_resident(self, scene, tile_offset=0, tile_stride=1):
        p, objs, nobj, cam = scenes.to_abi(scene)
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_render_resident(self._h, C.byref(p), objs, nobj, C.byref(cam), tile_offset, tile_stride,
                                          C.byref(stats)))
        return stats.as_dict()

    def upload_pixels(self, pixels):
        assert pixels.flags.c_contiguous and pixels.dtype == np.float64
        _check(lib().rtgr_upload_pixels(self._h, pixels.ctypes.data, pixels.shape[0]))

    def trace_resident(self, params, objs_arr, n_objs):
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_trace_resident(self._h, C.byref(params), objs_arr, n_objs, C.byref(stats)))
        return stats.as_dict()

    def make_canvas(self, params, cam):
        px = np.empty((cam.ni * cam.nj, 11))
        _check(lib().rtgr_make_canvas(self._h, C.byref(params), C.byref(cam), px.ctypes.data))
        return px

    def rhs_batch(self, params, states):
        states = np.ascontiguousarray(states, dtype=np.float64)
        out = np.empty_like(states)
        _check(lib().rtgr_rhs_batch(self._h, C.byref(params), _dp(states), states.shape[0], _dp(out)))
        return out

    # -- user-supplied metrics (the reference takes any callable metric, src:483) --------------------
    def compile_metric(self, source, par=()):
        """rtgr_metric_compile: CUDA C++ source of `rtgr_user_metric<T>` -> metric id for rtgr_params.metric."""
        mid = C.c_int32(-1)
        _check(lib().rtgr_metric_compile(self._h, source.encode() if isinstance(source, str) else source, C.byref(mid)))
        if len(par):
            self.set_metric_params(mid.value, par)
        return mid.value

    def set_metric_params(self, metric_id, par):
        arr = (C.c_double * max(1, len(par)))(*[float(v) for v in par])
        _check(lib().rtgr_metric_set_params(self._h, metric_id, arr, len(par)))

    def release_metric(self, metric_id):
        _check(lib().rtgr_metric_release(self._h, metric_id))

    def fp64_peak(self, dev_index=0, mode=1):
        tf, mhz = C.c_double(), C.c_double()
        _check(lib().rtgr_fp64_microbench(self._h, dev_index, mode, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value


class Frame:
    """rtgr_frame: one frame shared by several GPUs -- the cross-GPU dynamic tile queue.  The owner creates
    it (`Frame(ctx, ni, nj)`) and hands `handle` (64 bytes) to the other processes, which open it
    (`Frame(ctx, ni, nj, handle=...)`); every participant then calls `render(scene)` once per frame, with a
    barrier of the caller's between frames and before `read()`.  See include/raytracegr_cuda.h."""

    def __init__(self, ctx, ni, nj, handle=None):
        self._ctx, self.ni, self.nj = ctx, int(ni), int(nj)
        self._h = C.c_void_p()
        if handle is None:
            hb = (C.c_uint8 * _abi.RTGR_IPC_HANDLE_BYTES)()
            _check(lib().rtgr_frame_create(ctx._h, self.ni, self.nj, C.byref(self._h), hb))
            self.handle, self.owner = bytes(hb), True
        else:
            handle = bytes(handle)
            if len(handle) != _abi.RTGR_IPC_HANDLE_BYTES:
                raise RtgrError("an IPC handle has %d bytes" % _abi.RTGR_IPC_HANDLE_BYTES)
            hb = (C.c_uint8 * _abi.RTGR_IPC_HANDLE_BYTES).from_buffer_copy(handle)
            _check(lib().rtgr_frame_open(ctx._h, hb, self.ni, self.nj, C.byref(self._h)))
            self.handle, self.owner = handle, False
        if not hasattr(ctx, "_frames"):
            ctx._frames = []
        ctx._frames.append(self)

    def render(self, scene):
        """rtgr_render_frame: this participant's share of the frame; returns its stats."""
        p, objs, nobj, cam = scenes.to_abi(scene)
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_render_frame(self._h, C.byref(p), objs, nobj, C.byref(cam), C.byref(stats)))
        return stats.as_dict()

    def trace_canvas(self, params, objs_arr, n_objs, pixels):
        """rtgr_trace_canvas_frame: trace_rays on ONE page-locked Pixel canvas shared by all participants
        (`pixels`: (nj, ni, 11) float64 in memory every participant has mapped and page-locked, e.g. a
        SharedCanvas); this participant's share, rgb written in place.  Returns its stats."""
        assert pixels.dtype == np.float64 and pixels.flags.c_contiguous and pixels.shape == (self.nj, self.ni, 11)
        stats = _abi.rtgr_stats()
        _check(lib().rtgr_trace_canvas_frame(self._h, C.byref(params), objs_arr, n_objs, pixels.ctypes.data,
                                             self.ni, self.nj, C.byref(stats)))
        return stats.as_dict()

    def set_participants(self, n):
        """rtgr_frame_set_participants: tell the library how many GPUs share the frame (a tuning hint)."""
        _check(lib().rtgr_frame_set_participants(self._h, int(n)))

    def read(self):
        img = np.empty((self.nj, self.ni, 3), dtype=np.uint8)
        _check(lib().rtgr_frame_read(self._h, _u8p(img)))
        return img

    def clear(self):
        _check(lib().rtgr_frame_clear(self._h))

    def close(self):
        if self._h:
            lib().rtgr_frame_close(self._h)
            self._h = C.c_void_p()
            if self in getattr(self._ctx, "_frames", ()):
                self._ctx._frames.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _u8p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


_DEFAULT_CTX = None


def default_context():
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context([0])
    return _DEFAULT_CTX


def _params_of(metric):
    return _abi.default_params(metric.kind, M=metric.M, a=metric.a, r_formula=metric.r_formula)


def screen_widths(view_angle_deg, ni, nj, x_dir=(0, 1, 0, 0), y_dir=(0, 0, 0, 1)):
    """The `widthx`, `widthy` of a screen with a VERTICAL view angle of `view_angle_deg` and square pixels, for a unit
    `normal` (the reference's README speaks of "a screen with a certain width and height, with a view angle"; its
    make_canvas, src:458-478, takes the two width vectors): a pixel's ray direction is normal + dx widthx + dy widthy
    with dx, dy in (-1/2, 1/2), so |widthy| = 2 tan(angle/2) and |widthx| = |widthy| ni/nj.  90 degrees at 16:9 gives
    the (0, 32/9, 0, 0), (0, 0, 0, 2) of BASELINE configs[3]."""
    import math
    h = 2.0 * math.tan(math.radians(float(view_angle_deg)) / 2.0)
    w = h * float(ni) / float(nj)
    return tuple(w * float(v) for v in x_dir), tuple(h * float(v) for v in y_dir)


def make_canvas(metric, pos, widthx, widthy, normal, ni, nj, ctx=None):
    """make_canvas(metric, pos, widthx, widthy, normal, ni, nj) -> Canvas (src:458-478)."""
    ctx = ctx or default_context()
    cam = _abi.rtgr_camera()
    cam.pos[:] = [float(v) for v in pos]
    cam.widthx[:] = [float(v) for v in widthx]
    cam.widthy[:] = [float(v) for v in widthy]
    cam.normal[:] = [float(v) for v in normal]
    cam.ni, cam.nj = int(ni), int(nj)
    px = ctx.make_canvas(_params_of(metric), cam)
    return Canvas(px.reshape(nj, ni, 11))


def trace_rays(metric, objs, c, ctx=None):
    """trace_rays(metric, objs, c::Canvas)::Canvas (src:483-536): pure -- the input canvas is left
    untouched and a new one with the rgb fields filled in is returned."""
    ctx = ctx or default_context()
    arr = _marshal_objects(objs)
    # the new canvas lives in page-locked memory, which the kernel reads and writes in place
    buf = PinnedArray((c.nj, c.ni, 11))
    buf.array[...] = c.pixels
    ctx.trace_canvas(_params_of(metric), arr, len(objs), buf.array)
    return Canvas(buf.array, owner=buf)


def write_png(path, img8):
    """Minimal 8-bit RGB PNG writer (no third-party dependency)."""
    img8 = np.ascontiguousarray(img8, dtype=np.uint8)
    h, w, _ = img8.shape
    raw = b"".join(b"\x00" + img8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


outdir = "scenes"   # src:540


def _example(scene, filename, ctx=None):
    metric = Metric(scene.metric, scene.M, scene.a, scene.r_formula)
    objs = [Plane(o[1]) if o[0] == "plane" else Sphere(o[1], o[2], o[3]) for o in scene.objects]
    canvas = make_canvas(metric, scene.pos, scene.widthx, scene.widthy, scene.normal, scene.ni, scene.nj, ctx=ctx)
    canvas = trace_rays(metric, objs, canvas, ctx=ctx)
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, filename)
    if os.path.exists(path):
        os.remove(path)
    print('Output file is "%s"' % path)
    write_png(path, canvas.image8())
    return canvas


def example1(ctx=None):
    """src:542-576: flat-space sphere, writes scenes/sphere.png."""
    return _example(scenes.example1(), "sphere.png", ctx)


def example2(ctx=None):
    """src:578-612: sphere near a black hole, writes scenes/sphere2.png."""
    return _example(scenes.example2(), "sphere2.png", ctx)


def render_scene(scene, ctx=None, **kw):
    """Fused production path (device-side make_canvas + trace) for a scenes.Scene."""
    return (ctx or default_context()).render(scene, **kw)
