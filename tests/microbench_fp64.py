"""Prints the two FP64 DFMA microbenchmarks (pipe limit vs three-register-operand limit)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
ctx = pkg.Context([0])
for n in (1, 3):
    vals = [ctx.fp64_peak(0, n)[0] for _ in range(3)]
    print("fp64 DFMA microbench, %d register operand(s): best %.2f TFLOP/s (runs %s)" % (n, max(vals), ["%.2f" % v for v in vals]))
ctx.close()
