"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every
symbol the public header declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import conftest

ROOT = conftest.ROOT


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    header = open(os.path.join(ROOT, "include", "raytracegr_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rtgr_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(declared) == sorted(pkg._lib.SYMBOLS)
    assert L.rtgr_version() == 100


def test_library_contains_sm100a_code(pkg):
    out = subprocess.run(["cuobjdump", "--list-elf", pkg.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_default_params_are_the_reference_values(pkg):
    p = pkg._abi.rtgr_params()
    pkg.lib().rtgr_default_params(C.byref(p), pkg._abi.RTGR_KERR_SCHILD)
    assert (p.metric, p.r_formula, p.M, p.a) == (1, 0, 1.0, 0.0)                 # src:275-276
    assert (p.lambda0, p.lambda1) == (0.0, 100.0)                                 # src:497
    assert p.reltol == p.abstol == float(np.finfo(np.float64).eps) ** 0.75        # src:485
    assert p.hit_threshold == 0.01 and p.interp_points == 10 and p.maxiters == 100000
    q = pkg._abi.default_params(pkg._abi.RTGR_KERR_SCHILD)
    assert bytes(p) == bytes(q)


def test_struct_layouts(pkg):
    A = pkg._abi
    assert C.sizeof(A.rtgr_object) == 88
    assert C.sizeof(A.rtgr_camera) == 136
    assert C.sizeof(A.rtgr_params) == 72
    assert C.sizeof(A.rtgr_stats) == 56


@pytest.mark.skipif(conftest.has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError) as e:
        pkg.Context([0])
    assert "no CPU fallback" in str(e.value)


def test_scene_marshalling(pkg):
    sc = pkg.scenes.example2()
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    assert nobj == 3
    assert objs[0].kind == pkg._abi.RTGR_SPHERE and objs[0].radius == -10.0      # caelum, src:582
    assert objs[1].kind == pkg._abi.RTGR_PLANE and objs[1].time == -20.0         # frustum, src:583
    assert list(objs[2].pos) == [0.0, 4.0, 0.0, 0.0] and objs[2].radius == 0.5   # sphere, src:584-585
    assert list(cam.pos) == [0.0, 4.0, -2.0, 0.0] and (cam.ni, cam.nj) == (200, 200)
    c4 = pkg.scenes.config4()
    assert (c4.ni, c4.nj, c4.a) == (3840, 2160, 0.99)


def test_png_writer_roundtrip(pkg, tmp_path):
    img = (np.arange(7 * 5 * 3) % 256).astype(np.uint8).reshape(5, 7, 3)
    path = str(tmp_path / "x.png")
    pkg.write_png(path, img)
    from PIL import Image
    assert np.array_equal(np.array(Image.open(path)), img)


def test_host_objects_reject_abstract(pkg):
    from raytracegr_jl_b200 import host
    with pytest.raises(host.RtgrError):
        host._marshal_objects([object()])
    arr = host._marshal_objects([pkg.Sphere((0, 0, 0, 0), (1, 0, 0, 0), -10), pkg.Plane(-20)])
    assert arr[0].radius == -10 and arr[1].time == -20


def test_frame_entry_points_reject_null_arguments(pkg):
    """The shared-frame entry points validate their arguments before touching CUDA (so this runs without a GPU)."""
    L = pkg.lib()
    h = C.c_void_p()
    hb = (C.c_uint8 * pkg._abi.RTGR_IPC_HANDLE_BYTES)()
    assert L.rtgr_frame_create(None, 64, 64, C.byref(h), hb) != 0 and not h
    assert "NULL" in pkg._lib.last_error()
    assert L.rtgr_frame_open(None, hb, 64, 64, C.byref(h)) != 0 and not h
    assert L.rtgr_render_frame(None, None, None, 0, None, None) != 0
    assert "frame" in pkg._lib.last_error()
    assert L.rtgr_trace_canvas_frame(None, None, None, 0, None, 8, 8, None) != 0
    assert "frame" in pkg._lib.last_error()
    assert L.rtgr_frame_set_participants(None, 8) != 0
    assert L.rtgr_frame_read(None, None) != 0
    assert L.rtgr_frame_clear(None) != 0
    L.rtgr_frame_close(None)      # a no-op


def test_header_is_plain_c_and_links_from_c(pkg, tmp_path):
    """include/raytracegr_cuda.h compiles as pedantic C99, and a plain-C program (examples/c_abi_example.c) links
    against the library and calls it.  Without a GPU it must stop at rtgr_create with the no-fallback message
    (exit status 2); on a GPU box it renders example2 and must report all 40 000 rays."""
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    src = os.path.join(ROOT, "examples", "c_abi_example.c")
    exe = str(tmp_path / "c_abi_example")
    csrc = os.path.dirname(pkg.library_path())
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           src, "-L", csrc, "-lraytracegr_cuda", "-o", exe])
    env = dict(os.environ, LD_LIBRARY_PATH=csrc + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert "libraytracegr_cuda version 100" in out.stdout
    if conftest.has_gpu():
        assert out.returncode == 0, out.stderr
        assert "40000 rays" in out.stdout
    else:
        assert out.returncode == 2
        assert "no CPU fallback" in out.stderr


def test_baseline_configurations_match_the_survey_table(pkg):
    """the five BASELINE.json configurations, literal by literal."""
    S = pkg.scenes
    e1, e2, c3, c4, c5 = S.example1(), S.example2(), S.config3(), S.config4(), S.config5()
    assert (e1.metric, e1.ni, e1.nj, e1.pos) == (0, 200, 200, (0, 0, -2, 0))                      # src:552-557
    assert (e2.metric, e2.a, e2.ni, e2.nj, e2.pos) == (1, 0.0, 200, 200, (0, 4, -2, 0))             # src:588-593
    for s in (e1, e2, c3, c4, c5):
        assert s.M == 1.0 and s.r_formula == 0 and s.normal == (0, 0, 1, 0)
        assert s.objects[0] == ("sphere", (0, 0, 0, 0), (1, 0, 0, 0), -10.0) and s.objects[1] == ("plane", -20.0)
    assert (c3.a, c3.ni, c3.nj, c3.widthx, c3.widthy) == (0.9, 1920, 1080, (0, 16.0 / 9.0, 0, 0), (0, 0, 0, 1))
    assert (c4.a, c4.ni, c4.nj, c4.widthx, c4.widthy) == (0.99, 3840, 2160, (0, 32.0 / 9.0, 0, 0), (0, 0, 0, 2))
    assert (c5.a, c5.ni, c5.nj, c5.widthx, c5.widthy) == (0.9, 7680, 4320, (0, 16.0 / 9.0, 0, 0), (0, 0, 0, 1))
    for s in (e1, e2, c3, c4):
        assert s.tol == float(np.finfo(np.float64).eps) ** 0.75
    for s in (c3, c4, c5):       # square pixels
        assert abs(s.widthx[1] / s.ni - s.widthy[3] / s.nj) < 1e-15


def test_screen_widths_view_angle(pkg):
    # 90 degrees vertical at 16:9 is the camera of BASELINE configs[3] (scenes.config4)
    wx, wy = pkg.screen_widths(90.0, 3840, 2160)
    s = pkg.scenes.config4()
    assert max(abs(a - b) for a, b in zip(wx, s.widthx)) < 1e-14
    assert max(abs(a - b) for a, b in zip(wy, s.widthy)) < 1e-14
    # the reference's examples: unit widths at 1:1 are a view angle of 2 atan(1/2)
    import math
    wx, wy = pkg.screen_widths(math.degrees(2 * math.atan(0.5)), 200, 200)
    assert abs(wx[1] - 1.0) < 1e-14 and abs(wy[3] - 1.0) < 1e-14
