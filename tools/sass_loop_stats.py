#!/usr/bin/env python3
"""Developer tool: static statistics of the step loop of the trace kernels in a built library (no GPU needed):
size of the outermost loop in bytes (the loop has to fit the instruction caches), FP64 instructions inside it, DFMAs
with three distinct register sources (3 instead of 2 dispatch cycles unless one operand comes from the reuse cache)
and the non-FP64 rest.  The executed path per step attempt is a subset of the loop: tools/hotpath.py merges an ncu
source capture for that.  usage: sass_loop_stats.py [lib.so] [kernel-name substring ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raytracegr.jl_b200", "csrc", "libraytracegr_cuda.so")
want = sys.argv[2:] or ["trace_kernelILi1ELi0", "trace_stage_kernelILi1ELi0", "trace_kernelILi1ELi2", "trace_kernelILi0ELi0"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.split("\n"):
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
        funcs[cur] = []
    elif cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))


def op(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]


for name, ins in funcs.items():
    if not any(w in name for w in want):
        continue
    back = []
    for a, t in ins:
        m = re.search(r"BRA (0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            back.append((int(m.group(1), 16), a))
    if not back:
        continue
    lo, hi = max(back, key=lambda b: b[1] - b[0])
    other, fp64, n3, n3r = collections.Counter(), collections.Counter(), 0, 0
    for a, t in ins:
        if not lo <= a <= hi:
            continue
        o = op(t)
        if o in ("DFMA", "DMUL", "DADD", "DSETP"):
            fp64[o] += 1
            if o == "DFMA":
                srcs = [p.strip() for p in re.sub(r"^DFMA\s+", "", re.sub(r"^@!?U?P\d+\s+", "", t)).split(",")][1:4]
                regs = [re.sub(r"[-|]|\.reuse", "", s) for s in srcs if re.match(r"^-?\|?R\d+", s)]
                if len(regs) == 3 and len(set(regs)) == 3:
                    n3 += 1
                    n3r += any(".reuse" in s for s in srcs)
        else:
            other[o] += 1
    short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_raytracegr_cuda_cu_[0-9a-f]+\d\d", "", name)
    print("%s: %d instructions; step loop %#x..%#x = %d bytes; inside it FP64 %d %s, three-register DFMAs %d (%d with a .reuse "
          "operand), other %d %s" % (short, len(ins), lo, hi, hi - lo, sum(fp64.values()), dict(fp64), n3, n3r, sum(other.values()),
                                      dict(other.most_common(10))))
