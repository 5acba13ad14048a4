"""The scenes of the reference and of BASELINE.json, as plain data.

example1 / example2 restate the literals of src/RayTraceGR.jl:542-558 and :578-594;
configs 3-5 are the BASELINE.json / SURVEY.md section 8(d) extensions (same objects and camera
position as example2, spin and screen changed).
"""
from dataclasses import dataclass, field
from typing import List, Tuple

from . import _abi


@dataclass
class Scene:
    name: str
    metric: int
    M: float
    a: float
    objects: List[Tuple]           # ("sphere", pos4, vel4, radius) | ("plane", time)
    pos: Tuple[float, ...]
    widthx: Tuple[float, ...]
    widthy: Tuple[float, ...]
    normal: Tuple[float, ...]
    ni: int
    nj: int
    tol: float = _abi.REFERENCE_TOL
    r_formula: int = _abi.RTGR_R_AS_WRITTEN
    note: str = ""

    def with_size(self, ni, nj):
        from dataclasses import replace
        return replace(self, ni=int(ni), nj=int(nj))


def _objs(sphere_pos):
    # caelum (sky, radius -10), frustum (plane t = -20), sphere (radius 1/2): src:546-549 / :582-585
    return [("sphere", (0, 0, 0, 0), (1, 0, 0, 0), -10.0),
            ("plane", -20.0),
            ("sphere", sphere_pos, (1, 0, 0, 0), 0.5)]


def example1(ni=200, nj=200):
    return Scene("example1", _abi.RTGR_MINKOWSKI, 1.0, 0.0, _objs((0, 0, 0, 0)),
                 (0, 0, -2, 0), (0, 1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), ni, nj,
                 note="flat Minkowski sphere scene, reference default resolution (src:542-558)")


def example2(ni=200, nj=200, a=0.0):
    return Scene("example2", _abi.RTGR_KERR_SCHILD, 1.0, a, _objs((0, 4, 0, 0)),
                 (0, 4, -2, 0), (0, 1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), ni, nj,
                 note="sphere near the Kerr-Schild hole, M=1 a=0 as in src:275-276 (src:578-594)")


def config3(ni=1920, nj=1080):
    s = example2(ni, nj, a=0.9)
    s.name = "ks_a0.9_1080p"
    s.widthx = (0, 16.0 / 9.0, 0, 0)
    s.note = "Kerr-Schild a=0.9, 1920x1080, square pixels (BASELINE.json configs[2])"
    return s


def config4(ni=3840, nj=2160):
    s = example2(ni, nj, a=0.99)
    s.name = "ks_a0.99_4k_wide"
    s.widthx = (0, 32.0 / 9.0, 0, 0)
    s.widthy = (0, 0, 0, 2.0)
    s.note = "Kerr-Schild a=0.99, 3840x2160, 90 degree vertical field of view (BASELINE.json configs[3])"
    return s


def config5(ni=7680, nj=4320, tol=1e-8):
    s = config3(ni, nj)
    s.name = "ks_a0.9_8k_tol%g" % tol
    s.tol = tol
    s.note = "Kerr-Schild a=0.9, 7680x4320, tolerance sweep member (BASELINE.json configs[4])"
    return s


BY_NAME = {"example1": example1, "example2": example2, "config3": config3, "config4": config4,
           "config5": config5}


def to_abi(scene):
    """(params, objects array, n_objs, camera) ctypes values for a Scene."""
    p = _abi.default_params(scene.metric, M=scene.M, a=scene.a, r_formula=scene.r_formula, tol=scene.tol)
    arr = (_abi.rtgr_object * len(scene.objects))()
    for i, o in enumerate(scene.objects):
        if o[0] == "plane":
            arr[i].kind = _abi.RTGR_PLANE
            arr[i].time = float(o[1])
        else:
            arr[i].kind = _abi.RTGR_SPHERE
            arr[i].pos[:] = [float(v) for v in o[1]]
            arr[i].vel[:] = [float(v) for v in o[2]]
            arr[i].radius = float(o[3])
    cam = _abi.rtgr_camera()
    cam.pos[:] = [float(v) for v in scene.pos]
    cam.widthx[:] = [float(v) for v in scene.widthx]
    cam.widthy[:] = [float(v) for v in scene.widthy]
    cam.normal[:] = [float(v) for v in scene.normal]
    cam.ni, cam.nj = int(scene.ni), int(scene.nj)
    return p, arr, len(scene.objects), cam
