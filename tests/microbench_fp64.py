"""Prints the FP64 operand-mix microbenchmarks (see rtgr_fp64_microbench in include/raytracegr_cuda.h)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
ctx = pkg.Context([0])
NAMES = {1: "DFMA a=a*C1+C2 (1 reg)", 2: "DFMA a=a*b+C (2 reg)", 3: "DFMA a=b*c+a (3 reg)", 4: "DMUL a=a*b (2 reg)",
         5: "DADD a=a+b (2 reg)", 6: "DMUL a=a*C (1 reg)", 7: "DFMA a=b*b+a (2 distinct reg)",
         8: "DFMA a=b*c+a, b shared by neighbours", 9: "alternating 3-reg DFMA / 2-reg DMUL",
         10: "2-reg DFMA + 1 integer instr each", 11: "2-reg DFMA + 3 integer instr each",
         12: "3-reg DFMA + 1 integer instr each"}
base = None
for mode in sorted(NAMES):
    vals = [ctx.fp64_peak(0, mode)[0] for _ in range(3)]
    base = base or max(vals)
    print("fp64 microbench mode %d %-40s best %.2f DFMA-equivalent TFLOP/s (%.1f %% of mode 1)"
          % (mode, NAMES[mode], max(vals), 100 * max(vals) / base))
ctx.close()
