"""Static cross-check of julia/RayTraceGRCUDA.jl (never executed: no Julia in the image) and of the
`ccall` stub in INTEGRATION.md against include/raytracegr_cuda.h: every ccall names a declared symbol,
passes as many arguments as the C prototype takes, pointer arguments are Ptr/Ref/Cstring and scalars are
scalars of the right width, and the isbits mirrors of the ABI structs have the C sizes."""
import os
import re

import conftest

ROOT = conftest.ROOT


def _split_top(s):
    """split at commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def c_prototypes():
    h = open(os.path.join(ROOT, "include", "raytracegr_cuda.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b([a-z_0-9]+\s*\*?)\s*\b(rtgr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else _split_top(args)
        protos[name] = (ret, params)
    return protos


C_SCALARS = {"int": {"Cint", "Int32"}, "int32_t": {"Cint", "Int32"}, "int64_t": {"Int64", "Clonglong"},
             "uint64_t": {"UInt64", "Culonglong"}, "double": {"Float64", "Cdouble"}}


def _matching_paren(src, i):
    depth = 0
    for k in range(i, len(src)):
        if src[k] == "(":
            depth += 1
        elif src[k] == ")":
            depth -= 1
            if depth == 0:
                return k
    raise AssertionError("unbalanced ccall")


def julia_ccalls(src):
    calls = []
    for m in re.finditer(r"ccall\(", src):
        j = _matching_paren(src, m.end() - 1)
        parts = _split_top(src[m.end():j])
        sym = re.match(r"\(:(\w+),\s*\w+\)", parts[0])
        assert sym, parts[0]
        types = parts[2].strip()
        assert types.startswith("(") and types.endswith(")"), types
        calls.append((sym.group(1), parts[1], _split_top(types[1:-1]), parts[3:]))
    return calls


def _check_source(src, protos):
    calls = julia_ccalls(src)
    assert calls
    for name, ret, types, values in calls:
        assert name in protos, "ccall of an undeclared symbol: " + name
        cret, cparams = protos[name]
        assert len(types) == len(cparams) == len(values), (name, types, cparams, values)
        if cret == "int":
            assert ret == "Cint", (name, ret)
        elif cret == "void":
            assert ret == "Cvoid", (name, ret)
        for jt, cp in zip(types, cparams):
            if "*" in cp:
                assert jt.startswith(("Ptr{", "Ref{")) or jt == "Cstring", (name, jt, cp)
                if "char" in cp:
                    assert jt in ("Cstring", "Ptr{UInt8}", "Ptr{Cchar}"), (name, jt, cp)
            else:
                ctype = cp.split()[-2] if len(cp.split()) > 1 else cp
                assert ctype in C_SCALARS, (name, cp)
                assert jt in C_SCALARS[ctype], (name, jt, cp)
    return {c[0] for c in calls}


def test_julia_module_matches_the_header():
    protos = c_prototypes()
    assert len(protos) >= 30
    src = open(os.path.join(ROOT, "julia", "RayTraceGRCUDA.jl")).read()
    used = _check_source(src, protos)
    # the reference-facing entry points are all bound
    for need in ("rtgr_create", "rtgr_destroy", "rtgr_last_error", "rtgr_make_canvas", "rtgr_trace_canvas", "rtgr_render",
                 "rtgr_host_register", "rtgr_host_unregister", "rtgr_metric_compile", "rtgr_frame_create",
                 "rtgr_frame_open", "rtgr_render_frame", "rtgr_frame_read", "rtgr_frame_close"):
        assert need in used, need


def test_integration_md_stub_matches_the_header():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```julia\n(.*?)```", md, flags=re.S)
    assert blocks
    used = set()
    for b in blocks:
        if "ccall(" in b:
            used |= _check_source(b, c_prototypes())
    assert "rtgr_trace_canvas" in used


JL_SIZES = {"Int32": 4, "UInt32": 4, "Float64": 8, "UInt64": 8, "Int64": 8, "NTuple{4,Float64}": 32, "NTuple{3,T}": 24,
            "NTuple{D,T}": 32}


def test_julia_struct_mirrors_have_the_c_sizes(pkg):
    import ctypes as C
    src = open(os.path.join(ROOT, "julia", "RayTraceGRCUDA.jl")).read()
    want = {"CObject": C.sizeof(pkg._abi.rtgr_object), "CParams": C.sizeof(pkg._abi.rtgr_params),
            "CCamera": C.sizeof(pkg._abi.rtgr_camera), "Stats": C.sizeof(pkg._abi.rtgr_stats), "Pixel{T}": 88}
    for name, size in want.items():
        m = re.search(r"struct %s(?=[\s<])[^\n]*\n(.*?)\nend" % re.escape(name), src, flags=re.S)
        assert m, name
        fields = [l.split("#")[0].strip() for l in m.group(1).split("\n")]
        types = [f.split("::")[1].strip() for f in fields if "::" in f]
        total, align = 0, 1
        for t in types:                       # natural alignment, as Julia lays out isbits structs
            a = 4 if t in ("Int32", "UInt32") else 8
            total = (total + a - 1) // a * a + JL_SIZES[t]
            align = max(align, a)
        total = (total + align - 1) // align * align
        assert total == size, (name, total, size)


def test_runtests_jl_uses_only_names_the_module_exports():
    """julia/runtests.jl cannot be run here either: at least every module name it calls is exported or qualified."""
    mod = open(os.path.join(ROOT, "julia", "RayTraceGRCUDA.jl")).read()
    tests = open(os.path.join(ROOT, "julia", "runtests.jl")).read()
    exported = set(n.strip() for n in re.search(r"^export (.*?)\n\n", mod, flags=re.S | re.M).group(1).replace("\n", " ").split(","))
    for name in ("example1", "example2", "make_canvas", "trace_rays", "trace_rays!", "kerr_schild", "Object", "Sphere", "Plane", "screen_widths"):
        assert name in exported, name
        assert name.rstrip("!") in tests
    for qualified in set(re.findall(r"RayTraceGRCUDA\.(\w+!?)", tests)) - {"jl"}:      # ("RayTraceGRCUDA.jl" is the file name)
        assert re.search(r"^(function |const |struct |mutable struct |)%s\b" % re.escape(qualified), mod, flags=re.M) or \
            re.search(r"^%s\(" % re.escape(qualified), mod, flags=re.M), qualified
    assert os.path.exists(os.path.join(ROOT, "tests", "golden", "sphere.npy")) and os.path.exists(os.path.join(ROOT, "tests", "golden", "sphere2.npy"))
