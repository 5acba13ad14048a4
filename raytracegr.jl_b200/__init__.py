"""raytracegr.jl_b200 -- B200-native geodesic ray tracing hot path of RayTraceGR.jl.

The directory name carries a dot, so it is loaded under the module name
``raytracegr_jl_b200`` (see ``__graft_entry__.load_package``).

Contents
  csrc/      CUDA kernels (sm_100a) and the C ABI of include/raytracegr_cuda.h
  _abi.py    ctypes mirror of the ABI structs
  _lib.py    loader for csrc/libraytracegr_cuda.so (fails loudly if absent; no CPU fallback)
  host.py    host-side mirror of the reference's scene API (Sphere, Plane, make_canvas,
             trace_rays, example1, example2) on top of the C ABI
  scenes.py  the reference's example scenes and the BASELINE.json configurations
  metrics/   example user metrics for rtgr_metric_compile (CUDA C++ source compiled at run time)
"""
from . import _abi, scenes  # noqa: F401
from ._lib import lib, library_path, build_library  # noqa: F401
from .host import (  # noqa: F401
    Canvas, Context, Frame, PinnedArray, Plane, SharedCanvas, Sphere, example1, example2, kerr_schild, make_canvas, minkowski,
    render_scene, screen_widths, trace_rays, write_png, user_metric, check_metric_source, METRIC_SOURCES,
)
