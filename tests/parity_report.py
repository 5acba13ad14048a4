#!/usr/bin/env python3
"""Prints the parity numbers the north star names (object-id agreement, final position / momentum error,
RGB error) for the CUDA path against the oracle and the golden images -- the same comparisons the tests
assert, reported as numbers."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import oracle_lib  # noqa: E402
import parity  # noqa: E402

pkg = entry.load_package()
ctx = pkg.Context([0])
for name, size in (("example1", (200, 200)), ("example2", (200, 200)), ("config3", (192, 108)), ("config4", (192, 108))):
    sc = pkg.scenes.BY_NAME[name](ni=size[0], nj=size[1])
    p, objs, nobj, cam = pkg.scenes.to_abi(sc)
    px = oracle_lib.make_canvas(p, cam)
    ref = oracle_lib.trace_pixels(p, objs, nobj, px)
    mine = np.array(px, copy=True)
    out = ctx.trace_pixels(p, objs, nobj, mine, want=("final_state", "obj_id", "status", "nsteps"))
    res = parity.compare(ref, out, ref["pixels"][:, 8:], mine[:, 8:])
    line = {"scene": sc.name, "ni": sc.ni, "nj": sc.nj, **{k: res[k] for k in ("n", "id_agree", "n_id_mismatch", "max_ex", "max_eu", "rgb_max", "n_state_bad", "n_rgb_bad")},
            "steps_oracle": int(ref["stats"]["steps_accepted"]), "steps_gpu": int(out["stats"]["steps_accepted"])}
    g = os.path.join(ROOT, "tests", "golden", {"example1": "sphere.npy", "example2": "sphere2.npy"}.get(name, "none"))
    if os.path.exists(g):
        golden = np.load(g)
        img = np.rint(255.0 * np.clip(mine[:, 8:], 0, 1)).astype(np.uint8).reshape(sc.nj, sc.ni, 3)
        line["golden_png_pixels_bit_exact"] = float((img == golden).all(axis=2).mean())
    print(json.dumps(line), flush=True)
ctx.close()
