"""Prints the FP64 latency / parallelism probe (modes >= 100 of rtgr_fp64_microbench): DFMA warp-instructions
per cycle and scheduler for `chains` independent chains per thread and `w` warps per scheduler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
ctx = pkg.Context([0])
for lc, chains in enumerate((1, 2, 4, 8)):
    for w in (1, 2, 3, 4, 6, 8):
        tf, mhz = max(ctx.fp64_peak(0, 100 + 10 * lc + w) for _ in range(2))
        rate = tf * 1e12 / (2 * 32 * 4 * 148 * mhz * 1e6)     # warp-DFMA per cycle per scheduler
        print("chains %d warps/scheduler %d: %.2f TFLOP/s = %.3f warp-DFMA/cycle/scheduler (%.2f cycles per DFMA; in flight %d)"
              % (chains, w, tf, rate, 1.0 / rate, chains * w))
ctx.close()
