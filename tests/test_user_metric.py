"""User-supplied metrics: the reference's trace_rays takes ANY callable metric
(src:483) and differentiates through it with its Dual type (src:298-331).  Here the metric is CUDA
C++ source compiled at run time; these tests pin that path to the oracle's generic Dual -> dmetric ->
christoffel -> geodesic evaluation and to the reference's golden image."""
import os

import numpy as np
import pytest

import parity

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SCHWARZSCHILD_ISOTROPIC = 1000   # oracle/rtgr_oracle.cpp: test metric, m = params.M


def src(pkg, name):
    """A shipped metric source; `name + ":stationary"` prepends the author's declaration that g does not depend on x[0]."""
    name, _, decl = name.partition(":")
    text = open(os.path.join(pkg.METRIC_SOURCES, name + ".cu")).read()
    return ("#pragma rtgr stationary\n" + text) if decl == "stationary" else text


# the reference's kerr_schild as a user metric: the 4x4 matrix with 4 partials, the same declared stationary (3
# partials), and in Kerr-Schild form (f and k alone: closed-form right-hand side, no matrix inverse)
KS_SOURCES = ["kerr_schild_as_written", "kerr_schild_as_written:stationary", "kerr_schild_form"]


def random_states(n, seed=1):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-8, 8, (3 * n, 3))
    xyz = xyz[np.linalg.norm(xyz, axis=1) >= 1.5][:n]
    st = np.zeros((n, 8))
    st[:, 0] = rng.uniform(-20, 0, n)
    st[:, 1:4] = xyz
    st[:, 4:] = rng.uniform(-1, 1, (n, 4))
    return st


# ------------------------------- no GPU needed ---------------------------------------------------
@pytest.mark.parametrize("name", KS_SOURCES + ["schwarzschild_isotropic", "schwarzschild_isotropic:stationary"])
def test_shipped_metric_sources_compile_for_sm100a(pkg, name):
    # NVRTC cross-compiles without a device: the embedded device headers + the user function build
    log = pkg.check_metric_source(src(pkg, name))
    assert "error" not in log.lower() and "warning" not in log.lower(), log


def test_compile_error_is_reported_with_the_users_line(pkg):
    bad = "template <class T> __device__ void rtgr_user_metric(const T x[4], T g[4][4], const double* par) {\n  g[0][0] = undefined_symbol;\n}\n"
    with pytest.raises(pkg.host.RtgrError) as e:
        pkg.check_metric_source(bad)
    assert "undefined_symbol" in str(e.value) and "user_metric(2)" in str(e.value)


def test_oracle_isotropic_metric_is_a_vacuum_solution_far_field(pkg, oracle):
    # sanity of the oracle's extra test metric: g -> eta at infinity, g_tt = 0 on the horizon rho = m/2,
    # dmetric agrees with central differences
    p = pkg._abi.default_params(ORACLE_SCHWARZSCHILD_ISOTROPIC, M=1.0)
    g = oracle.metric(p, [0.0, 3e8, 0.0, 0.0])
    assert np.allclose(g, np.diag([-1.0, 1, 1, 1]), atol=1e-7)
    assert abs(oracle.metric(p, [0.0, 0.3, 0.4, 0.0])[0, 0]) < 1e-30
    x = np.array([-3.0, 2.0, -1.5, 0.7])
    g0, dg = oracle.dmetric(p, x)
    for c in range(4):
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[c] += h; xm[c] -= h
        fd = (oracle.metric(p, xp) - oracle.metric(p, xm)) / (2 * h)
        assert np.allclose(dg[:, :, c], fd, atol=1e-8)


# ------------------------------- on the GPU ------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", KS_SOURCES)
@pytest.mark.parametrize("a", [0.0, 0.9])
def test_user_kerr_schild_rhs_and_canvas_match_oracle(pkg, oracle, ctx, a, name):
    # the reference's own kerr_schild handed in as a USER metric: the generic path (seeded duals ->
    # g, dg -> inverse -> contraction; or f, k -> closed form) against the oracle's, which follows the reference
    # line by line
    A = pkg._abi
    mid = ctx.compile_metric(src(pkg, name), par=(1.0, a))
    try:
        pu = A.default_params(mid)
        po = A.default_params(A.RTGR_KERR_SCHILD, a=a)
        st = random_states(200000)
        out = ctx.rhs_batch(pu, st)
        ref = oracle.rhs_batch(po, st)
        assert np.array_equal(out[:, :4], st[:, 4:])
        scale = np.abs(ref[:, 4:]).max(axis=1, keepdims=True)
        rel = np.abs(out[:, 4:] - ref[:, 4:]) / scale
        assert np.median(rel) < 1e-15 and np.quantile(rel, 0.999) < 1e-12, (np.median(rel), rel.max())
        # ... and against the hand-specialised built-in kernel
        blt = ctx.rhs_batch(po, st)
        rel2 = np.abs(out[:, 4:] - blt[:, 4:]) / scale
        assert np.quantile(rel2, 0.999) < 1e-12
        sc = pkg.scenes.example2(ni=40, nj=30)
        _, _, _, cam = pkg.scenes.to_abi(sc)
        px = ctx.make_canvas(pu, cam)
        pxo = oracle.make_canvas(po, cam)
        assert np.abs(px[:, :8] - pxo[:, :8]).max() < 1e-13
    finally:
        ctx.release_metric(mid)


@pytest.mark.gpu
@pytest.mark.parametrize("name", KS_SOURCES)
def test_user_kerr_schild_reproduces_the_golden_image(pkg, ctx, name):
    # example2 rendered THROUGH THE GENERIC PATH reproduces the reference's shipped sphere2.png
    A = pkg._abi
    golden = np.load(os.path.join(HERE, "golden", "sphere2.npy"))
    mid = ctx.compile_metric(src(pkg, name), par=(1.0, 0.0))
    try:
        sc = pkg.scenes.example2()
        p, objs, nobj, cam = pkg.scenes.to_abi(sc)
        pu = A.default_params(mid)
        buf = pkg.PinnedArray((sc.nj, sc.ni, 11))
        buf.array[...] = ctx.make_canvas(pu, cam).reshape(sc.nj, sc.ni, 11)
        out = ctx.trace_canvas(pu, objs, nobj, buf.array, want=("obj_id", "status"))
        img = np.rint(255.0 * np.clip(buf.array[:, :, 8:], 0, 1)).astype(np.uint8)
        assert (img == golden).all(axis=2).mean() >= 0.999
        ids, counts = np.unique(out["obj_id"], return_counts=True)
        assert dict(zip(ids.tolist(), counts.tolist())) == {1: 31338, 2: 5154, 3: 3508}   # SURVEY 4.2 census
        # the fused render path (device make_canvas with the user metric) gives the same picture
        from dataclasses import replace
        r = ctx.render(replace(sc, metric=mid), want=("rgb8",))
        assert (r["rgb8"] == img).all(axis=2).mean() >= 0.9999
        buf.free()
    finally:
        ctx.release_metric(mid)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["schwarzschild_isotropic", "schwarzschild_isotropic:stationary"])
def test_user_schwarzschild_isotropic_matches_oracle(pkg, oracle, ctx, name):
    # a metric the reference does not ship: rhs and traced rays against the oracle's generic evaluation
    A = pkg._abi
    m = 1.0
    mid = ctx.compile_metric(src(pkg, name), par=(m,))
    try:
        pu = A.default_params(mid)
        po = A.default_params(ORACLE_SCHWARZSCHILD_ISOTROPIC, M=m)
        st = random_states(100000, seed=2)
        out = ctx.rhs_batch(pu, st)
        ref = oracle.rhs_batch(po, st)
        scale = np.abs(ref[:, 4:]).max(axis=1, keepdims=True)
        rel = np.abs(out[:, 4:] - ref[:, 4:]) / scale
        assert np.median(rel) < 1e-15 and np.quantile(rel, 0.999) < 1e-12
        sc = pkg.scenes.example2(ni=64, nj=64)
        _, objs, nobj, cam = pkg.scenes.to_abi(sc)
        pxo = oracle.make_canvas(po, cam)
        px = ctx.make_canvas(pu, cam)
        assert np.abs(px[:, :8] - pxo[:, :8]).max() < 1e-13
        refr = oracle.trace_pixels(po, objs, nobj, pxo)
        mine = np.array(pxo, copy=True)
        got = ctx.trace_pixels(pu, objs, nobj, mine, want=("final_state", "obj_id", "status", "nsteps"))
        res = parity.compare(refr, got, refr["pixels"][:, 8:], mine[:, 8:])
        assert res["id_agree"] >= parity.ID_AGREEMENT_MIN, res
        assert res["n_state_bad"] <= 0.002 * res["n"], res
        assert res["n_rgb_bad"] <= 0.002 * res["n"], res
        assert len(set(got["obj_id"].tolist())) >= 2          # the picture is not trivial
    finally:
        ctx.release_metric(mid)


@pytest.mark.gpu
def test_user_metric_errors(pkg, ctx):
    A = pkg._abi
    with pytest.raises(pkg.host.RtgrError):
        ctx.compile_metric("this is not C++")
    p = A.default_params(A.RTGR_USER_METRIC_BASE + 7)     # never compiled
    with pytest.raises(pkg.host.RtgrError):
        ctx.rhs_batch(p, np.zeros((1, 8)))
