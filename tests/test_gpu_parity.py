"""Parity tests proper: the CUDA path, called through the C ABI (ctypes on libraytracegr_cuda.so),
against the CPU oracle on identical inputs and against the reference's golden images.

Bars (BASELINE.json north_star): terminating object id equal on >= 99.9 % of pixels, final position
and momentum within 1e-8 relative, RGB within 1/255.
"""
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def random_states(n, seed=0):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-8, 8, (3 * n, 3))
    xyz = xyz[np.linalg.norm(xyz, axis=1) >= 1.2][:n]
    st = np.zeros((n, 8))
    st[:, 0] = rng.uniform(-20, 0, n)
    st[:, 1:4] = xyz
    st[:, 4:] = rng.uniform(-1, 1, (n, 4))
    return st


def test_native_library_is_loaded(pkg, ctx):
    # the CUDA extension is in-tree and is the thing that runs
    maps = open("/proc/self/maps").read()
    assert "libraytracegr_cuda.so" in maps
    assert ctx.n_devices == 1


@pytest.mark.parametrize("a,rf", [(0.0, 0), (0.9, 0), (0.99, 0), (0.9, 1)])
def test_rhs_batch(pkg, oracle, ctx, a, rf):
    # 2^20 seeded random states + the reference's seven test points (test/runtests.jl:41-44)
    p = pkg._abi.default_params(pkg._abi.RTGR_KERR_SCHILD, a=a, r_formula=rf)
    st = random_states(1 << 20)
    seven = np.array([[0, 2.0 * (i & 1), 2.0 * (i & 2), 2.0 * (i & 4), -1.0, 0.3, -0.2, 0.5] for i in range(1, 8)])
    st[:7] = seven
    out = ctx.rhs_batch(p, st)
    ref = oracle.rhs_batch(p, st)
    truth = oracle.rhs_batch(p, st[:100000], extended=True)
    assert np.array_equal(out[:, :4], st[:, 4:])
    scale = np.abs(ref[:, 4:]).max(axis=1, keepdims=True)
    rel = np.abs(out[:, 4:] - ref[:, 4:]) / scale
    assert np.median(rel) < 1e-15
    assert np.quantile(rel, 0.999) < 1e-13
    # against the extended-precision truth the kernel is as accurate as the as-written evaluation
    ts = np.abs(truth[:, 4:]).max(axis=1, keepdims=True)
    e_k = (np.abs(out[:100000, 4:] - truth[:, 4:]) / ts).max()
    e_r = (np.abs(ref[:100000, 4:] - truth[:, 4:]) / ts).max()
    assert e_k < 1e-11 and e_k < 10 * e_r + 1e-13, (e_k, e_r)


def test_rhs_minkowski_is_free_motion(pkg, ctx):
    p = pkg._abi.default_params(pkg._abi.RTGR_MINKOWSKI)
    st = random_states(1000)
    out = ctx.rhs_batch(p, st)
    assert np.array_equal(out[:, :4], st[:, 4:]) and np.all(out[:, 4:] == 0)


@pytest.mark.parametrize("name", ["example1", "example2", "config3", "config4"])
def test_make_canvas(pkg, oracle, ctx, name):
    sc = pkg.scenes.BY_NAME[name](ni=97, nj=45)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    out = ctx.make_canvas(p, cam)
    ref = oracle.make_canvas(p, cam)
    if name == "example1":
        assert np.array_equal(out, ref)
    else:
        assert np.allclose(out, ref, rtol=4e-15, atol=1e-15)


@pytest.fixture(scope="module")
def examples(pkg, oracle, ctx):
    """example1 and example2 at the reference's resolution: oracle and CUDA results on the same canvas."""
    res = {}
    for name in ("example1", "example2"):
        sc = pkg.scenes.BY_NAME[name]()
        p, objs, nobj, cam = pkg.scenes.to_abi(sc)
        px = oracle.make_canvas(p, cam)
        ref = oracle.trace_pixels(p, objs, nobj, px)
        mine = np.array(px, copy=True)
        out = ctx.trace_pixels(p, objs, nobj, mine, want=("final_state", "obj_id", "status", "nsteps"))
        out["pixels"] = mine
        res[name] = (sc, ref, out)
    return res


def test_example2_parity_with_oracle(examples):
    sc, ref, out = examples["example2"]
    r = parity.compare(ref, out, ref["pixels"][:, 8:], out["pixels"][:, 8:])
    print("example2 parity:", r)
    assert r["id_agree"] >= parity.ID_AGREEMENT_MIN
    assert r["n_state_bad"] == 0, r          # every ray, horizon-hugging ones included, within 1e-8
    assert r["n_rgb_bad"] == 0, r
    assert np.array_equal(out["status"], ref["status"])
    # input canvas fields untouched, only rgb written (trace_rays is pure on pos/normal)
    assert np.array_equal(out["pixels"][:, :8], ref["pixels"][:, :8])
    s = out["stats"]
    assert s["rays"] == 40000 and s["steps_rejected"] == 0
    assert abs(s["steps_accepted"] - ref["stats"]["steps_accepted"]) < 1e-3 * s["steps_accepted"]
    assert s["rhs_evals"] == 6 * (s["steps_accepted"] + s["steps_rejected"]) + 2 * s["rays"]


def test_example2_reproduces_golden_image(examples):
    sc, ref, out = examples["example2"]
    gold = np.load(os.path.join(HERE, "golden", "sphere2.npy"))
    img = np.rint(255 * np.clip(out["pixels"][:, 8:], 0, 1)).astype(np.uint8).reshape(200, 200, 3)
    diff = np.abs(img.astype(int) - gold.astype(int)).max(axis=2)
    assert (diff == 0).mean() >= 0.999, (diff == 0).mean()
    assert (diff <= 1).mean() >= 0.9995
    census = np.bincount(out["obj_id"], minlength=4)
    assert abs(int(census[1]) - 31338) <= 5 and abs(int(census[2]) - 5154) <= 5 and abs(int(census[3]) - 3508) <= 5


def test_example1_parity_with_oracle(examples):
    sc, ref, out = examples["example1"]
    r = parity.compare(ref, out, ref["pixels"][:, 8:], out["pixels"][:, 8:])
    print("example1 parity:", r)
    assert r["id_agree"] >= parity.ID_AGREEMENT_MIN
    assert r["n_state_bad"] == 0 and r["n_rgb_bad"] == 0, r
    # mismatches (if any) are silhouette-edge rays
    bad = np.nonzero(ref["obj_id"] != out["obj_id"])[0]
    if len(bad):
        rad = np.hypot(bad % 200 - 99.5, bad // 200 - 99.5)
        assert rad.min() > 32.5 and rad.max() < 34.2


def test_example1_golden_image(examples):
    sc, ref, out = examples["example1"]
    gold = np.load(os.path.join(HERE, "golden", "sphere.npy"))
    img = np.rint(255 * np.clip(out["pixels"][:, 8:], 0, 1)).astype(np.uint8).reshape(200, 200, 3)
    eq = (img == gold).all(axis=2)
    assert eq.mean() >= 0.996          # the reference's own edge pixels are not reproducible across machines
    jj, ii = np.nonzero(~eq)
    rad = np.hypot(ii - 99.5, jj - 99.5)
    assert rad.min() > 32.5 and rad.max() < 34.2


@pytest.mark.parametrize("name,ni,nj", [("config3", 192, 108), ("config4", 192, 108)])
def test_spinning_configs_parity(pkg, oracle, ctx, name, ni, nj):
    sc = pkg.scenes.BY_NAME[name](ni=ni, nj=nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    ref = oracle.trace_pixels(p, objs, nobj, px)
    mine = np.array(px, copy=True)
    out = ctx.trace_pixels(p, objs, nobj, mine, want=("final_state", "obj_id", "status", "nsteps"))
    r = parity.compare(ref, out, ref["pixels"][:, 8:], mine[:, 8:])
    print(name, "parity:", r)
    assert r["id_agree"] >= parity.ID_AGREEMENT_MIN
    assert r["n_state_bad"] <= 1e-3 * r["n"], r
    assert r["n_rgb_bad"] <= 1e-3 * r["n"], r


@pytest.mark.parametrize("tol", [1e-6, 1e-8, 1e-10])
def test_tolerance_sweep(pkg, oracle, ctx, tol):
    sc = pkg.scenes.config5(ni=128, nj=72, tol=tol)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle.make_canvas(p, cam)
    ref = oracle.trace_pixels(p, objs, nobj, px)
    mine = np.array(px, copy=True)
    out = ctx.trace_pixels(p, objs, nobj, mine, want=("final_state", "obj_id", "status", "nsteps"))
    same = ref["obj_id"] == out["obj_id"]
    assert same.mean() >= 0.998
    ex, eu = parity.state_rel_err(ref["final_state"], out["final_state"])
    assert np.quantile(ex[same], 0.99) < 200 * tol and np.quantile(eu[same], 0.9) < 2000 * tol


def test_render_equals_canvas_plus_trace_and_tiles_partition(pkg, ctx):
    # ragged screen; the fused render == make_canvas + trace_pixels; interleaved tile subsets
    # (what N ranks do) reproduce the single pass bit for bit
    sc = pkg.scenes.example2(ni=150, nj=77)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    want = ("rgb8", "rgb_f64", "final_state", "obj_id", "status", "nsteps")
    full = ctx.render(sc, want=want)
    assert full["stats"]["rays"] == 150 * 77
    px = ctx.make_canvas(p, cam)
    tp = ctx.trace_pixels(p, objs, nobj, px, want=("final_state", "obj_id", "status", "nsteps"))
    assert np.array_equal(tp["final_state"], full["final_state"])
    assert np.array_equal(px[:, 8:], full["rgb_f64"])
    assert np.array_equal(np.rint(255 * np.clip(px[:, 8:], 0, 1)).astype(np.uint8).reshape(77, 150, 3), full["rgb8"])
    parts = None
    rays = 0
    for r in range(3):
        parts = ctx.render(sc, want=want, tile_offset=r, tile_stride=3, out=parts)
        rays += parts["stats"]["rays"]
    assert rays == 150 * 77
    for k in want:
        assert np.array_equal(parts[k], full[k]), k
    # determinism: a second run is bit-identical although rays are scheduled dynamically
    again = ctx.render(sc, want=want)
    for k in want:
        assert np.array_equal(again[k], full[k]), k


@pytest.mark.parametrize("name,ni,nj", [("config4", 256, 128), ("example2", 200, 200), ("example1", 328, 76),
                                        ("config4", 237, 131), ("config4", 40, 36)])
def test_rgb8_patch_staging(pkg, ctx, monkeypatch, name, ni, nj):
    """With the patch staging (the kernel variant used when the image lives in another GPU's memory; forced here by
    RTGR_RGB8_STAGING=1) the RGB8 image leaves the kernel as whole 8x4-pixel patches (96 bytes collected in shared
    memory, twelve 8-byte stores) when the row length is a multiple of 8, byte by byte otherwise and for the patches
    a border tile cuts: either way it is the quantised rgb_f64 of the same launch and equal to the image of the
    plain kernel, also for interleaved tile subsets written into one image (what the ranks of a shared frame do)."""
    sc = pkg.scenes.BY_NAME[name]().with_size(ni, nj)
    monkeypatch.setenv("RTGR_RGB8_STAGING", "0")
    plain = ctx.render(sc, want=("rgb8",))["rgb8"]
    monkeypatch.setenv("RTGR_RGB8_STAGING", "1")
    out = ctx.render(sc, want=("rgb8", "rgb_f64"))
    q = np.rint(255 * np.clip(out["rgb_f64"], 0, 1)).astype(np.uint8).reshape(nj, ni, 3)
    assert np.array_equal(out["rgb8"], q)
    only = ctx.render(sc, want=("rgb8",))          # the image alone (the production call)
    assert np.array_equal(only["rgb8"], q)
    parts = None
    for r in range(5):
        parts = ctx.render(sc, want=("rgb8",), tile_offset=r, tile_stride=5, out=parts)
    assert np.array_equal(parts["rgb8"], q)
    assert np.array_equal(plain, q)


@pytest.mark.parametrize("name,ni,nj", [("config4", 237, 131), ("example2", 200, 200), ("example1", 97, 64)])
def test_chunkwise_ray_reads_equal_ray_by_ray(pkg, ctx, monkeypatch, name, ni, nj):
    """Rays that come from a Pixel array are read a chunk (32 rays, 2816 bytes, coalesced) at a time into shared
    memory (trace_pixels_kernel); RTGR_CHUNK_RAYS=0 reads them ray by ray as the render kernel would.  Same results
    bit for bit: 1-D arrays whose length is not a multiple of 32, ragged canvases, tile subsets."""
    sc = pkg.scenes.BY_NAME[name]().with_size(ni, nj)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px0 = ctx.make_canvas(p, cam)
    want = ("final_state", "obj_id", "status", "nsteps")
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("RTGR_CHUNK_RAYS", flag)
        a = px0[: ni * nj - 13].copy()                      # 1-D, ragged end
        ra = ctx.trace_pixels(p, objs, nobj, a, want=want)
        c = px0.reshape(nj, ni, 11).copy()                  # canvas (pageable: staged through device memory)
        rc = ctx.trace_canvas(p, objs, nobj, c, want=want)
        halves = px0.reshape(nj, ni, 11).copy()             # two interleaved tile subsets of one canvas
        for r in range(2):
            ctx.trace_canvas(p, objs, nobj, halves, tile_offset=r, tile_stride=2)
        res[flag] = (a, ra, c, rc, halves)
    a0, ra0, c0, rc0, h0 = res["0"]
    a1, ra1, c1, rc1, h1 = res["1"]
    assert np.array_equal(a0, a1) and np.array_equal(c0, c1) and np.array_equal(h0, h1) and np.array_equal(c1, h1)
    for k in want:
        assert np.array_equal(ra0[k], ra1[k]), k
        assert np.array_equal(rc0[k], rc1[k]), k
    assert np.array_equal(c1.reshape(-1, 11)[: ni * nj - 13], a1)
    assert np.any(c1[:, :, 8:] != 0.0)


@pytest.mark.parametrize("pinned", [True, False])
def test_trace_canvas_in_place(pkg, ctx, pinned):
    # rtgr_trace_canvas (the trace_rays drop-in with the canvas shape): ragged screen, page-locked
    # (zero copy: the kernel reads/writes the host array in place) and pageable (staged) buffers,
    # whole canvas and interleaved tile subsets; everything bit-identical with the fused render
    sc = pkg.scenes.example2(ni=150, nj=77)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    want = ("final_state", "obj_id", "status", "nsteps")
    full = ctx.render(sc, want=want + ("rgb_f64",))
    canvas = ctx.make_canvas(p, cam).reshape(77, 150, 11)
    if pinned:
        buf = pkg.host.PinnedArray((77, 150, 11))
        px = buf.array
        assert pkg.lib().rtgr_host_is_pinned(px.ctypes.data) == 1
    else:
        px = np.empty((77, 150, 11))
        assert pkg.lib().rtgr_host_is_pinned(px.ctypes.data) == 0
    px[...] = canvas
    out = ctx.trace_canvas(p, objs, nobj, px, want=want)
    assert out["stats"]["rays"] == 150 * 77
    assert np.array_equal(px[:, :, :8], canvas[:, :, :8])            # pos/normal untouched
    assert np.array_equal(px.reshape(-1, 11)[:, 8:], full["rgb_f64"])
    for k in want:
        assert np.array_equal(out[k], full[k]), k
    # three "ranks" share one canvas: each traces its tiles in place
    px[...] = canvas
    rays = 0
    acc = {k: np.zeros_like(full[k]) for k in want}
    for r in range(3):
        o = ctx.trace_canvas(p, objs, nobj, px, tile_offset=r, tile_stride=3, want=want)
        rays += o["stats"]["rays"]
        if r == 0:   # only this rank's tiles are coloured so far
            assert 0 < np.count_nonzero(px[:, :, 10]) < 150 * 77
        for k in want:
            acc[k] += o[k]     # untouched entries are zero
    assert rays == 150 * 77
    assert np.array_equal(px.reshape(-1, 11)[:, 8:], full["rgb_f64"])
    for k in want:
        assert np.array_equal(acc[k], full[k]), k
    if pinned:
        buf.free()


def test_host_register_makes_a_buffer_zero_copy(pkg, ctx):
    # what the Julia wrapper does with the memory of an Array{Pixel{Float64},2}: pin it in place once
    sc = pkg.scenes.example1(ni=64, nj=40)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = ctx.make_canvas(p, cam).reshape(40, 64, 11).copy()
    ref = px.copy()
    ctx.trace_canvas(p, objs, nobj, ref)
    L = pkg.lib()
    assert L.rtgr_host_register(px.ctypes.data, px.nbytes) == 0
    try:
        assert L.rtgr_host_is_pinned(px.ctypes.data) == 1
        ctx.trace_canvas(p, objs, nobj, px)
    finally:
        assert L.rtgr_host_unregister(px.ctypes.data) == 0
    assert np.array_equal(px, ref)
    assert px[:, :, 10].max() > 0


def test_full_size_properties_1080p(pkg, oracle, ctx):
    # BASELINE config 3 at full size: size-independent properties instead of a full oracle run
    sc = pkg.scenes.config3()
    out = ctx.render(sc, want=("rgb8", "final_state", "obj_id", "status", "nsteps"))
    n = sc.ni * sc.nj
    st = out["stats"]
    assert st["rays"] == n
    assert st["rhs_evals"] == 6 * (st["steps_accepted"] + st["steps_rejected"]) + 2 * n
    assert np.all(out["status"] == 0)                      # every ray ends on an object
    assert np.all(out["obj_id"] >= 1)
    fs = out["final_state"]
    # rays end ON an object: |min_distance| tiny relative to the scale of the distance function
    t, x, y, z = fs[:, 0], fs[:, 1], fs[:, 2], fs[:, 3]
    d_sky = 100 - (x * x + y * y + z * z)
    d_pl = t + 20
    d_sp = (x - 4) ** 2 + y * y + z * z - 0.25
    dmin = np.minimum(np.minimum(d_sky, d_pl), d_sp)
    assert np.abs(dmin).max() < 1e-9
    # the null condition g(u,u) = 0 is conserved along every ray (checked on a sample with the oracle metric)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    idx = np.random.default_rng(0).integers(0, n, 300)
    for k in idx:
        g = oracle.metric(p, fs[k, :4])
        u = fs[k, 4:]
        assert abs(u @ g @ u) <= 1e-7 * np.abs(u).max() ** 2
    # ... and, with an independent numpy evaluation of the metric, on EVERY ray of the frame, together with the
    # Killing energy g_{0b} u^b between the ray's first and last state (the metric is stationary).  A wrong
    # Christoffel contraction violates both at O(1); the oracle's own rays keep them to 4e-12 / 1e-9 at 192x108.
    s0 = ctx.make_canvas(p, cam)[:, :8]
    null_rel, e_rel = parity.conservation_errors(s0, fs, sc.M, sc.a)
    assert null_rel.max() < 1e-8, null_rel.max()
    assert e_rel.max() < 1e-6, e_rel.max()
    # a lattice of ~50 000 rays agrees with the oracle to the north-star bars
    _assert_lattice_parity(pkg, oracle, sc, out, 41)


def _lattice_parity(pkg, oracle, sc, out, stride):
    """A lattice subsample of a full-size frame against the oracle; returns (sub, ref ids, ids equal, ex, eu, dRGB8)."""
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    n = sc.ni * sc.nj
    sub = np.arange(stride // 2, n, stride)
    px = oracle.make_canvas(p, cam)[sub]
    ref = oracle.trace_pixels(p, objs, nobj, px)
    same = ref["obj_id"] == out["obj_id"][sub]
    ex = eu = None
    if out.get("final_state") is not None:
        ex, eu = parity.state_rel_err(ref["final_state"], out["final_state"][sub])
    rgb_ref = np.rint(255 * np.clip(ref["pixels"][:, 8:], 0, 1)).astype(int)
    rgb_out = out["rgb8"].reshape(-1, 3)[sub].astype(int)
    return sub, ref, same, ex, eu, np.abs(rgb_ref - rgb_out).max(axis=1)


def _on_object_edge(obj_img, pix, ids):
    """True where pixel `pix` of the (nj, ni) id image has, within one pixel, every id of `ids[k]` (a pair): the ray
    grazes the boundary between the two objects in the image (a silhouette or the shadow edge)."""
    nj, ni = obj_img.shape
    jj, ii = pix // ni, pix % ni
    ok = np.zeros(len(pix), dtype=bool)
    for k in range(len(pix)):
        win = obj_img[max(0, jj[k] - 1):jj[k] + 2, max(0, ii[k] - 1):ii[k] + 2]
        ok[k] = all(int(v) in win for v in ids[k])
    return ok


def _assert_lattice_parity(pkg, oracle, sc, out, stride, rgb_tol=1, state=True):
    """The north-star bars (BASELINE.json) on a lattice of the full-size frame: terminating object id equal on
    >= 99.9 % of the rays, and EVERY id mismatch lies on the image boundary between the two objects in question (a
    ray grazing an object edge or the shadow edge); final position and momentum within 1e-8 relative and RGB within
    1/255 on >= 99.9 % of the id-agreeing rays."""
    sub, ref, same, ex, eu, drgb = _lattice_parity(pkg, oracle, sc, out, stride)
    assert len(sub) >= 50000 or stride == 1, len(sub)
    assert same.mean() >= parity.ID_AGREEMENT_MIN, (same.mean(), int((~same).sum()), len(sub))
    bad = np.flatnonzero(~same)
    if len(bad):
        img = out["obj_id"].reshape(sc.nj, sc.ni)
        pairs = np.stack([ref["obj_id"][bad], out["obj_id"][sub][bad]], axis=1)
        edge = _on_object_edge(img, sub[bad], pairs)
        assert edge.all(), ("id mismatches away from any object edge", sub[bad][~edge][:10], pairs[~edge][:10])
    if state and ex is not None:
        ok = (ex[same] < parity.STATE_RTOL) & (eu[same] < parity.STATE_RTOL)
        assert ok.mean() >= 0.999, (ok.mean(), float(ex[same].max()), float(eu[same].max()))
    assert (drgb[same] <= rgb_tol).mean() >= 0.999, float((drgb[same] <= rgb_tol).mean())
    return dict(rays=len(sub), id_agree=float(same.mean()), mismatches=int((~same).sum()),
                max_ex=float(ex[same].max()) if ex is not None else None,
                max_eu=float(eu[same].max()) if eu is not None else None)


def test_full_size_properties_4k_config4(pkg, oracle, ctx):
    # BASELINE configs[3], the bench workload, at its full 3840x2160: exact work identities, every ray
    # ends on an object, the frame does not depend on how it is cut into shards, and a lattice of
    # ~50 000 rays agrees with the oracle to the north-star bars
    sc = pkg.scenes.config4()
    n = sc.ni * sc.nj
    out = ctx.render(sc, want=("rgb8", "final_state", "obj_id", "status"))
    st = out["stats"]
    assert st["rays"] == n
    assert st["rhs_evals"] == 6 * (st["steps_accepted"] + st["steps_rejected"]) + 2 * n
    assert np.all(out["status"] == 0) and np.all(out["obj_id"] >= 1)
    fs = out["final_state"]
    t, x, y, z = fs[:, 0], fs[:, 1], fs[:, 2], fs[:, 3]
    dmin = np.minimum(np.minimum(100 - (x * x + y * y + z * z), t + 20), (x - 4) ** 2 + y * y + z * z - 0.25)
    assert np.abs(dmin).max() < 1e-9
    _assert_lattice_parity(pkg, oracle, sc, out, 163)
    # 5 interleaved shards (what 5 ranks would trace) give the same image bit for bit
    parts = None
    for r in range(5):
        parts = ctx.render(sc, want=("rgb8",), tile_offset=r, tile_stride=5, out=parts)
    assert np.array_equal(parts["rgb8"], out["rgb8"])


@pytest.mark.parametrize("tol", [1e-6, 1e-10])
def test_full_size_properties_8k_config5(pkg, oracle, ctx, tol):
    # BASELINE configs[4] at its full 7680x4320 (33 M rays), both ends of the tolerance sweep
    sc = pkg.scenes.config5(tol=tol)
    n = sc.ni * sc.nj
    out = ctx.render(sc, want=("rgb8", "obj_id", "status"))
    st = out["stats"]
    assert st["rays"] == n
    assert st["rhs_evals"] == 6 * (st["steps_accepted"] + st["steps_rejected"]) + 2 * n
    assert (st["steps_rejected"] > 0) == (tol > 1e-7)          # rejections appear only at loose tolerances
    assert np.all(out["status"] == 0) and np.all(out["obj_id"] >= 1)
    # ~50 000 rays against the oracle at the same tolerance.  At tol = 1e-6 two correct integrators differ by about
    # the tolerance itself in the end point, which the sphere's 12-fold colour pattern magnifies: the colour bar
    # is 3/255 there (1/255 at 1e-10); ids to the north-star bar at both ends.
    _assert_lattice_parity(pkg, oracle, sc, out, 661, rgb_tol=(1 if tol < 1e-8 else 3), state=False)


def test_edge_cases(pkg, ctx):
    A = pkg._abi
    p = A.default_params(A.RTGR_MINKOWSKI)
    objs = (A.rtgr_object * 1)()
    # n = 0
    r = ctx.trace_pixels(p, objs, 0, np.zeros((0, 11)))
    assert r["stats"]["rays"] == 0
    # no objects: the ray runs to lambda1 and is coloured red
    px = np.zeros((1, 11)); px[0, :8] = [0, 0, 0, 0, -1, 1, 0, 0]
    r = ctx.trace_pixels(p, objs, 0, px, want=("status", "obj_id", "final_state"))
    assert r["status"][0] == A.STATUS_LAMBDA_END and r["obj_id"][0] == 0
    assert list(px[0, 8:]) == [1.0, 0.0, 0.0]
    assert np.allclose(r["final_state"][0, :4], [-100, 100, 0, 0], rtol=1e-13)
    # NaN input is flagged
    px[0, 1] = np.nan
    r = ctx.trace_pixels(p, objs, 0, px, want=("status",))
    assert r["status"][0] == A.STATUS_NONFINITE
    # ragged ray counts around warp / block / 1024-block boundaries
    sc = pkg.scenes.example2(ni=64, nj=33)
    pp, oo, no, cam = pkg.scenes.to_abi(sc)
    canvas = ctx.make_canvas(pp, cam)
    base = np.array(canvas, copy=True)
    full = ctx.trace_pixels(pp, oo, no, base, want=("final_state",))
    for n in (1, 31, 33, 1023, 1025, 2047):
        sub = np.array(canvas[:n], copy=True)
        r = ctx.trace_pixels(pp, oo, no, sub, want=("final_state",))
        assert np.array_equal(r["final_state"], full["final_state"][:n])
        assert np.array_equal(sub[:, 8:], base[:n, 8:])
    # maxiters honoured
    pp.maxiters = 5
    one = np.array(canvas[:1], copy=True)
    r = ctx.trace_pixels(pp, oo, no, one, want=("status",))
    assert r["status"][0] == A.STATUS_MAXITERS and r["stats"]["steps_accepted"] + r["stats"]["steps_rejected"] == 5


@pytest.mark.parametrize("name,ni,nj", [("example2", 200, 200), ("config4", 384, 216)])
def test_fp64_controller_build_agrees(pkg, ctx, tmp_path, name, ni, nj):
    # The shipped kernel evaluates the PI controller's step-size FACTOR with lg2/ex2.approx in FP32 and scales the
    # error norm with a one-Newton reciprocal (accept/reject and all state arithmetic are FP64).  The side-by-side
    # build csrc/libraytracegr_cuda_ctl64.so (-DRTGR_CONTROLLER_FP64: FP64 log/exp, IEEE divisions -- the controller
    # of SURVEY A.2/A.3 to the letter) must give the same picture: ids 100 %, final states to 1e-10.
    import subprocess
    import sys
    lib64 = os.path.join(os.path.dirname(pkg._lib.library_path()), "libraytracegr_cuda_ctl64.so")
    assert os.path.exists(lib64), "run `make -C raytracegr.jl_b200/csrc` (builds both libraries)"
    dump = str(tmp_path / "ctl64.npz")
    env = dict(os.environ, RTGR_LIBRARY=lib64)
    subprocess.run([sys.executable, os.path.join(HERE, "render_dump.py"), name, str(ni), str(nj), dump], check=True, env=env, timeout=600)
    alt = np.load(dump)
    assert str(alt["library"]) == "libraytracegr_cuda_ctl64.so" and int(alt["loaded"]) == 1
    sc = pkg.scenes.BY_NAME[name](ni=ni, nj=nj)
    out = ctx.render(sc, want=("rgb8", "obj_id", "final_state", "status", "nsteps"))
    assert np.array_equal(out["obj_id"], alt["obj_id"])
    assert np.array_equal(out["status"], alt["status"])
    ex, eu = parity.state_rel_err(alt["final_state"], out["final_state"])
    assert ex.max() < 1e-10 and eu.max() < 1e-10, (ex.max(), eu.max())
    assert np.abs(out["rgb8"].astype(int) - alt["rgb8"].astype(int)).max() <= 1
    # the two controllers take (almost) the same steps: the factor differs by ~1e-7 relative
    att = out["stats"]["steps_accepted"] + out["stats"]["steps_rejected"]
    assert abs(att - int(alt["attempts"])) <= 2e-4 * att, (att, int(alt["attempts"]))


def test_ray_inside_rho_less_than_a_is_flagged(pkg, ctx):
    # rho < a under the as-written radius (src:284): sqrt of a negative number -- Julia would throw a DomainError and
    # abort the render; the kernel ends that ray with status NONFINITE and keeps going
    A = pkg._abi
    sc = pkg.scenes.config4(ni=8, nj=8)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = ctx.make_canvas(p, cam).copy()
    px[5, :4] = [0.0, 0.3, 0.2, 0.1]          # rho = 0.374 < a = 0.99
    px[5, 4:8] = [-1.0, 0.1, 0.2, 0.3]
    r = ctx.trace_pixels(p, objs, nobj, px, want=("status", "obj_id", "final_state"))
    assert r["status"][5] == A.STATUS_NONFINITE
    assert np.all(np.delete(r["status"], 5) == A.STATUS_EVENT)
    # the stopped ray is coloured at its last finite state like any other (src:513-533): the start point, a miss (red)
    assert np.array_equal(r["final_state"][5], px[5, :8]) and r["obj_id"][5] == 0
    assert list(px[5, 8:]) == [1.0, 0.0, 0.0]


def test_error_reporting(pkg, ctx):
    import ctypes as C
    A = pkg._abi
    from raytracegr_jl_b200.host import RtgrError
    p = A.default_params(7)                          # unknown metric
    objs = (A.rtgr_object * 1)()
    with pytest.raises(RtgrError, match="metric"):
        ctx.trace_pixels(p, objs, 0, np.zeros((1, 11)))
    p = A.default_params(A.RTGR_KERR_SCHILD)
    with pytest.raises(RtgrError, match="n_objs"):
        ctx.trace_pixels(p, objs, 17, np.zeros((1, 11)))
    p.reltol = -1.0
    with pytest.raises(RtgrError, match="toler"):
        ctx.trace_pixels(p, objs, 0, np.zeros((1, 11)))


def test_host_api_examples(pkg, ctx, tmp_path, monkeypatch):
    # the reference-facing entry points: example1()/example2() write scenes/sphere.png, scenes/sphere2.png
    monkeypatch.chdir(tmp_path)
    c2 = pkg.example2(ctx=ctx)
    assert os.path.exists(tmp_path / "scenes" / "sphere2.png")
    gold = np.load(os.path.join(HERE, "golden", "sphere2.npy"))
    from PIL import Image
    img = np.array(Image.open(tmp_path / "scenes" / "sphere2.png"))
    assert (np.abs(img.astype(int) - gold.astype(int)).max(axis=2) == 0).mean() >= 0.999
    # trace_rays is pure: a new canvas comes back, the input is untouched
    canvas = pkg.make_canvas(pkg.minkowski, (0, 0, -2, 0), (0, 1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0), 20, 20, ctx=ctx)
    before = canvas.pixels.copy()
    objs = [pkg.Sphere((0, 0, 0, 0), (1, 0, 0, 0), -10), pkg.Plane(-20), pkg.Sphere((0, 0, 0, 0), (1, 0, 0, 0), 0.5)]
    out = pkg.trace_rays(pkg.minkowski, objs, canvas, ctx=ctx)
    assert np.array_equal(canvas.pixels, before)
    assert out.rgb.max() > 0 and np.array_equal(out.pixels[:, :, :8], before[:, :, :8])


def test_multi_device_matches_single(pkg, ctx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sc = pkg.scenes.example2(ni=150, nj=77)
    want = ("rgb8", "rgb_f64", "final_state", "obj_id", "status", "nsteps")
    one = ctx.render(sc, want=want)
    with pkg.Context(list(range(min(4, torch.cuda.device_count())))) as multi:
        many = multi.render(sc, want=want)
        for k in want:
            assert np.array_equal(one[k], many[k]), k
        p, objs, nobj, cam = pkg.scenes.to_abi(sc)
        px1 = ctx.make_canvas(p, cam)
        px2 = np.array(px1, copy=True)
        a = ctx.trace_pixels(p, objs, nobj, px1, want=("final_state", "obj_id"))
        b = multi.trace_pixels(p, objs, nobj, px2, want=("final_state", "obj_id"))
        assert np.array_equal(px1, px2) and np.array_equal(a["final_state"], b["final_state"])
        assert np.array_equal(a["obj_id"], b["obj_id"])
        # the in-place canvas drop-in over several devices: page-locked (every device reads and writes
        # the one host array directly) and pageable (staged per device, tiles gathered on the host)
        for pinned in (True, False):
            buf = pkg.PinnedArray((77, 150, 11)) if pinned else None
            pxc = buf.array if pinned else np.empty((77, 150, 11))
            pxc[...] = ctx.make_canvas(p, cam).reshape(77, 150, 11)
            c = multi.trace_canvas(p, objs, nobj, pxc, want=("final_state", "obj_id"))
            assert np.array_equal(pxc.reshape(-1, 11), px1), pinned
            assert np.array_equal(c["final_state"], a["final_state"]) and np.array_equal(c["obj_id"], a["obj_id"])
            if buf is not None:
                buf.free()


def test_ray_paths(pkg, oracle, ctx):
    # rtgr_trace_paths: every accepted step of a ray, what the reference's solve stores
    # with save_everystep although trace_rays only reads sol[end]
    A = pkg._abi
    # Minkowski: straight lines x(lambda) = x0 + lambda u0, the last point is the event state
    sc = pkg.scenes.example1(ni=16, nj=12)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = ctx.make_canvas(p, cam)
    ref = ctx.trace_pixels(p, objs, nobj, np.array(px, copy=True), want=("final_state", "obj_id", "status", "nsteps"))
    r = ctx.trace_paths(p, objs, nobj, px[:, :8], max_points=64)
    assert np.array_equal(r["final_state"], ref["final_state"]) and np.array_equal(r["obj_id"], ref["obj_id"])
    assert np.array_equal(r["npoints"], ref["nsteps"] + 1)           # initial state + one point per accepted step
    for i in range(px.shape[0]):
        k = r["npoints"][i]
        pts = r["paths"][i, :k]
        assert np.array_equal(pts[0], np.concatenate([[0.0], px[i, :8]]))
        assert np.all(np.diff(pts[:, 0]) > 0)
        assert np.allclose(pts[:, 1:5], px[i, :4] + pts[:, :1] * px[i, 4:8], rtol=0, atol=1e-12)
        assert np.array_equal(pts[-1, 1:], ref["final_state"][i])
        assert np.all(r["paths"][i, k:] == 0)
    # Kerr-Schild a = 0.9: along every path the ray stays null and keeps its Killing energy u_t = g_tb u^b
    sc = pkg.scenes.config3(ni=12, nj=8)
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = ctx.make_canvas(p, cam)
    ref = ctx.trace_pixels(p, objs, nobj, np.array(px, copy=True), want=("final_state", "nsteps"))
    r = ctx.trace_paths(p, objs, nobj, px[:, :8], max_points=4096)
    assert np.array_equal(r["final_state"], ref["final_state"])
    assert np.array_equal(r["npoints"], ref["nsteps"] + 1)
    for i in range(0, px.shape[0], 7):
        k = r["npoints"][i]
        pts = r["paths"][i, :k]
        assert np.array_equal(pts[-1, 1:], ref["final_state"][i])
        e0 = None
        for q in pts[:: max(1, k // 40)]:
            g = oracle.metric(p, q[1:5])
            u = q[5:9]
            assert abs(u @ g @ u) <= 1e-8 * np.abs(u).max() ** 2
            e = g[0] @ u
            e0 = e if e0 is None else e0
            assert abs(e - e0) <= 1e-8 * abs(e0)
    # truncation: the first max_points-1 points and the last one are kept, npoints still counts them all
    t = ctx.trace_paths(p, objs, nobj, px[:, :8], max_points=16)
    assert np.array_equal(t["npoints"], r["npoints"])
    for i in range(px.shape[0]):
        assert r["npoints"][i] > 16
        assert np.array_equal(t["paths"][i, :15], r["paths"][i, :15])
        assert np.array_equal(t["paths"][i, 15], r["paths"][i, r["npoints"][i] - 1])
    # ... also for a ray that ends WITHOUT an event (here: lambda1 reached in an empty Kerr-Schild scene)
    p_end = A.default_params(A.RTGR_KERR_SCHILD, a=0.9)
    p_end.lambda1 = 6.0
    none = (A.rtgr_object * 1)()
    full = ctx.trace_paths(p_end, none, 0, px[:, :8], max_points=4096)
    cut = ctx.trace_paths(p_end, none, 0, px[:, :8], max_points=8)
    assert np.array_equal(cut["npoints"], full["npoints"]) and np.array_equal(cut["status"], full["status"])
    # (rays captured by the hole stall at the quasi-horizon until dt < dtmin; the others reach lambda1)
    ended = np.flatnonzero((full["status"] == A.STATUS_LAMBDA_END) & (full["npoints"] > 8))
    assert len(ended) >= 1 and set(np.unique(full["status"])) <= {A.STATUS_LAMBDA_END, A.STATUS_DT_MIN}
    for i in np.flatnonzero(full["npoints"] > 8):
        assert np.array_equal(cut["paths"][i, :7], full["paths"][i, :7])
        assert np.array_equal(cut["paths"][i, 7], full["paths"][i, min(full["npoints"][i], 4096) - 1])
    for i in ended:
        assert cut["paths"][i, 7, 0] == 6.0                       # the last point is at lambda1
    # the run-time compiled user-metric kernel records the same paths as the built-in one
    src = open(os.path.join(pkg.METRIC_SOURCES, "kerr_schild_as_written.cu")).read()
    mid = ctx.compile_metric(src, par=(1.0, 0.9))
    try:
        uu = ctx.trace_paths(A.default_params(mid), objs, nobj, px[:, :8], max_points=4096)
        same = np.abs(uu["npoints"] - r["npoints"]) <= 2
        assert same.mean() >= 0.95
        ex, eu = parity.state_rel_err(r["final_state"], uu["final_state"])
        assert np.quantile(ex, 0.95) < 1e-8
    finally:
        ctx.release_metric(mid)
