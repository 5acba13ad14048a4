"""Test helper (run as a subprocess): renders a named scene with the CUDA library RTGR_LIBRARY points at and dumps
obj_id / final_state / rgb8 / work counters to an .npz -- how the test-suite runs a SECOND build of the library
(e.g. csrc/libraytracegr_cuda_ctl64.so) beside the one loaded in the test process.
usage: render_dump.py <scene name> <ni> <nj> <out.npz>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

name, ni, nj, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
pkg = entry.load_package()
ctx = pkg.Context([0])
sc = pkg.scenes.BY_NAME[name](ni=ni, nj=nj)
out = ctx.render(sc, want=("rgb8", "obj_id", "final_state", "status", "nsteps"))
maps = open("/proc/self/maps").read()
np.savez(path, rgb8=out["rgb8"], obj_id=out["obj_id"], final_state=out["final_state"], status=out["status"],
         nsteps=out["nsteps"], attempts=out["stats"]["steps_accepted"] + out["stats"]["steps_rejected"],
         library=os.path.basename(pkg._lib.library_path()), loaded=int(os.path.basename(pkg._lib.library_path()) in maps))
ctx.close()
