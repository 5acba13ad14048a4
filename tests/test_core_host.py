"""The product's per-lane arithmetic (csrc/rtgr_core.cuh + rtgr_trace.cuh), compiled for the host by
tests/host_shim.cpp, against the oracle.  Runs without a GPU; the same comparisons run on the real
kernels in test_gpu_parity.py."""
import numpy as np
import pytest

import parity


def random_states(n, seed=0):
    # x in U[-8,8]^3 rejected if rho < 1.2, t in U[-20,0], u in U[-1,1]^4
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-8, 8, (3 * n, 3))
    xyz = xyz[np.linalg.norm(xyz, axis=1) >= 1.2][:n]
    st = np.zeros((n, 8))
    st[:, 0] = rng.uniform(-20, 0, n)
    st[:, 1:4] = xyz
    st[:, 4:] = rng.uniform(-1, 1, (n, 4))
    return st


@pytest.mark.parametrize("a,rf", [(0.0, 0), (0.9, 0), (0.99, 0), (0.9, 1), (0.0, 1)])
def test_lean_rhs_matches_as_written_rhs(pkg, oracle, shim, a, rf):
    p = pkg._abi.default_params(pkg._abi.RTGR_KERR_SCHILD, a=a, r_formula=rf)
    st = random_states(50000)
    ref = oracle.rhs_batch(p, st)
    truth = oracle.rhs_batch(p, st, extended=True)
    out = shim.rhs_batch(p, st)
    assert np.array_equal(out[:, :4], st[:, 4:])                 # xdot = u (src:360)
    scale = np.abs(truth[:, 4:]).max(axis=1, keepdims=True)
    err_lean = (np.abs(out[:, 4:] - truth[:, 4:]) / scale).max()
    err_ref = (np.abs(ref[:, 4:] - truth[:, 4:]) / scale).max()
    assert err_lean < 1e-11, err_lean
    assert err_lean < 10 * err_ref + 1e-13                        # as accurate as the as-written evaluation
    assert np.median(np.abs(out[:, 4:] - ref[:, 4:]) / scale) < 1e-15


def test_rhs_seven_reference_points(pkg, oracle, shim):
    # the points of test/runtests.jl:41-44
    p = pkg._abi.default_params(pkg._abi.RTGR_KERR_SCHILD)
    st = np.array([[0, 2.0 * (i & 1), 2.0 * (i & 2), 2.0 * (i & 4), -1.0, 0.3, -0.2, 0.5] for i in range(1, 8)])
    ref = oracle.rhs_batch(p, st)
    out = shim.rhs_batch(p, st)
    assert np.allclose(out, ref, rtol=1e-13, atol=1e-15)


def test_rhs_outside_domain_is_nan(pkg, oracle, shim):
    # rho < a under the as-written radius: sqrt of a negative number (Julia would throw)
    p = pkg._abi.default_params(pkg._abi.RTGR_KERR_SCHILD, a=0.9)
    st = np.array([[0, 0.3, 0.2, 0.1, -1, 0.1, 0.2, 0.3]], dtype=float)
    assert np.isnan(oracle.rhs_batch(p, st)[0, 4:]).all()
    assert np.isnan(shim.rhs_batch(p, st)[0, 4:]).all()


@pytest.mark.parametrize("name", ["example1", "example2", "config3", "config4"])
def test_canvas_matches_oracle(pkg, oracle, shim, name):
    sc = pkg.scenes.BY_NAME[name](ni=33, nj=21)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    ref = oracle.make_canvas(p, cam)
    out = shim.make_canvas(p, cam)
    if name == "example1":
        assert np.array_equal(ref, out)          # Minkowski canvas is bit-identical by construction
    else:
        assert np.allclose(ref, out, rtol=4e-15, atol=1e-15)
    # the initial 4-velocity is null and past-directed (src:472-474)
    for k in range(0, ref.shape[0], 37):
        g = oracle.metric(p, out[k, :4])
        u = out[k, 4:8]
        assert abs(u @ g @ u) < 1e-13
        assert u[0] < 0


def test_example1_trace_is_step_for_step_identical(pkg, oracle, shim):
    sc = pkg.scenes.example1()
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    ref = oracle.trace_pixels(p, objs, nobj, px)
    out = shim.trace_pixels(p, objs, nobj, px)
    assert np.array_equal(ref["obj_id"], out["obj_id"])
    assert np.array_equal(ref["nsteps"], out["nsteps"])
    assert np.array_equal(ref["status"], out["status"])
    ex, eu = parity.state_rel_err(ref["final_state"], out["final_state"])
    assert ex.max() < 1e-13 and eu.max() == 0.0
    assert np.abs(out["rgb"] - ref["pixels"][:, 8:]).max() < 1e-11
    assert out["counters"]["accepted"] == ref["stats"]["steps_accepted"]


@pytest.mark.parametrize("name,ni,nj", [("example2", 80, 80), ("config3", 96, 54), ("config4", 96, 54)])
def test_kerr_schild_trace_parity(pkg, oracle, shim, name, ni, nj):
    sc = pkg.scenes.BY_NAME[name](ni=ni, nj=nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    ref = oracle.trace_pixels(p, objs, nobj, px)
    out = shim.trace_pixels(p, objs, nobj, px)
    res = parity.compare(ref, out, ref["pixels"][:, 8:], out["rgb"])
    assert res["id_agree"] >= parity.ID_AGREEMENT_MIN, res
    assert res["n_state_bad"] <= 0.001 * res["n"], res
    assert res["n_rgb_bad"] <= 0.001 * res["n"], res
    # same work: 6 RHS per attempt; step counts agree to rounding-level differences in dt
    assert abs(out["counters"]["accepted"] - ref["stats"]["steps_accepted"]) <= 2e-3 * ref["stats"]["steps_accepted"]


@pytest.mark.parametrize("tol", [1e-6, 1e-8, 1e-10])
def test_tolerance_sweep_parity(pkg, oracle, shim, tol):
    sc = pkg.scenes.config5(ni=64, nj=36, tol=tol)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    ref = oracle.trace_pixels(p, objs, nobj, px)
    out = shim.trace_pixels(p, objs, nobj, px)
    res = parity.compare(ref, out, ref["pixels"][:, 8:], out["rgb"])
    assert res["id_agree"] >= 0.995, res     # 2304 rays: allow a couple of edge rays
    # at loose tolerances the two evaluations may choose different step sequences; both stay within
    # (a modest multiple of) the requested tolerance of each other
    ex, eu = parity.state_rel_err(ref["final_state"], out["final_state"])
    same = ref["obj_id"] == out["obj_id"]
    assert np.quantile(ex[same], 0.99) < 200 * tol and np.quantile(eu[same], 0.9) < 2000 * tol


def test_render_tiles_cover_the_screen_once(pkg, shim):
    # ragged screen (not a multiple of the 32x32 tile): every pixel traced exactly once, and the
    # union of interleaved tile subsets (what N ranks / N devices do) equals the single pass
    sc = pkg.scenes.example2(ni=70, nj=45)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    full = shim.render_tiles(p, objs, nobj, cam)
    assert full["counters"]["rays"] == 70 * 45
    assert np.all(full["obj_id"] > 0)
    parts = None
    rays = 0
    for r in range(3):
        parts = shim.render_tiles(p, objs, nobj, cam, tile_offset=r, tile_stride=3, out=parts)
        rays += parts["counters"]["rays"]
    assert rays == 70 * 45
    for k in ("rgb8", "rgb", "final_state", "obj_id", "status", "nsteps"):
        assert np.array_equal(full[k], parts[k]), k
    # render == make_canvas + trace_pixels
    px = shim.make_canvas(p, cam)
    tp = shim.trace_pixels(p, objs, nobj, px)
    assert np.array_equal(tp["final_state"], full["final_state"])
    assert np.array_equal(tp["rgb"], full["rgb"])
    img = np.rint(255 * np.clip(tp["rgb"], 0, 1)).astype(np.uint8).reshape(45, 70, 3)
    assert np.array_equal(img, full["rgb8"])


def test_edge_cases(pkg, shim):
    A = pkg._abi
    # no objects: nothing to hit, the ray runs to lambda1 and is coloured red (src:527-528)
    p = A.default_params(A.RTGR_MINKOWSKI)
    objs = (A.rtgr_object * 1)()
    px = np.zeros((1, 11)); px[0, :8] = [0, 0, 0, 0, -1, 1, 0, 0]
    r = shim.trace_pixels(p, objs, 0, px)
    assert r["status"][0] == A.STATUS_LAMBDA_END and r["obj_id"][0] == 0
    assert list(r["rgb"][0]) == [1.0, 0.0, 0.0]
    assert np.allclose(r["final_state"][0, :4], [-100, 100, 0, 0], rtol=1e-13)
    # NaN initial state is reported, not propagated silently
    px[0, 1] = np.nan
    r = shim.trace_pixels(p, objs, 0, px)
    assert r["status"][0] == A.STATUS_NONFINITE
    # zero rays
    r = shim.trace_pixels(p, objs, 0, np.zeros((0, 11)))
    assert r["counters"]["rays"] == 0
    # a ray that dives below rho = a under the as-written radius stops with NONFINITE
    sc = pkg.scenes.example2(ni=1, nj=1, a=0.9)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = np.zeros((1, 11)); px[0, :8] = [0, 0.5, 0.1, 0.0, -1, 0.2, 0.1, 0.0]
    r = shim.trace_pixels(p, objs, nobj, px)
    assert r["status"][0] in (A.STATUS_NONFINITE, A.STATUS_EVENT)
    # maxiters is honoured
    sc = pkg.scenes.example2(ni=1, nj=1)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    p.maxiters = 5
    px = shim.make_canvas(p, cam)
    r = shim.trace_pixels(p, objs, nobj, px)
    assert r["status"][0] == A.STATUS_MAXITERS and r["counters"]["attempts"] == 5


@pytest.mark.parametrize("name,ni,nj", [("example2", 60, 60), ("config4", 64, 36)])
def test_conservation_laws_along_rays(pkg, oracle, shim, name, ni, nj):
    """What a wrong Christoffel symbol cannot fake ("index order is invisible in flat space"): along every
    traced ray the null norm g(u,u) = 0 and, the metric being stationary, the Killing energy g_{0b} u^b are conserved
    from the first to the last state -- for the oracle AND for the product's arithmetic, through hundreds of steps and a
    blueshift of up to 3e4 in |u|."""
    sc = pkg.scenes.BY_NAME[name](ni=ni, nj=nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    for res in (oracle.trace_pixels(p, objs, nobj, px), shim.trace_pixels(p, objs, nobj, px)):
        s0, s1 = px[:, :8], res["final_state"]
        assert np.isfinite(s1).all()
        null_rel, e_rel = parity.conservation_errors(s0, s1, sc.M, sc.a)
        assert null_rel.max() < 1e-9
        assert e_rel.max() < 1e-8
        assert np.abs(s1[:, 4:]).max() > 1e3          # the sample does contain strongly blueshifted (captured) rays
