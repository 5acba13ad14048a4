#!/usr/bin/env python3
"""Developer tool: static FP64 instruction mix of trace_kernel<KERR_SCHILD, AS_WRITTEN>'s hot loop, read
from `cuobjdump -sass` of the in-tree library (no GPU needed).  The per-attempt estimate assumes the
rolled stage loop runs 6 times and one of the 6 stage-state variants runs per iteration."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "raytracegr.jl_b200", "csrc", "libraytracegr_cuda.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern = sys.argv[1] if len(sys.argv) > 1 else "trace_kernelILi1ELi0"
ins, on = [], False
for l in txt.split("\n"):
    if "Function :" in l:
        on = kern in l
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
back = []
for a, t in ins:
    m = re.search(r"BRA (0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((int(m.group(1), 16), a))
outer = max((b for b in back if b[1] - b[0] > 0x3000 and b[1] < 0x9000), key=lambda b: b[1] - b[0])
stage = max((b for b in back if outer[0] < b[0] and b[1] < outer[1]), key=lambda b: b[1] - b[0])
# the RHS starts where the stage-state switch variants join: the most common forward BRA target inside the stage loop
tg = collections.Counter()
for a, t in ins:
    m = re.search(r"^BRA (0x[0-9a-f]+)", t)
    if m and stage[0] <= a < stage[1]:
        tg[int(m.group(1), 16)] += 1
rhs0 = tg.most_common(1)[0][0]
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "F2F", "I2F", "F2I", "FRND", "MUFU")


def mix(lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a < hi:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
            c[op] += 1
    return c


def show(name, lo, hi, scale=1.0):
    c = mix(lo, hi)
    f = {k: c[k] for k in FP64 if c[k]}
    n64 = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DSETP"]
    print("%-22s %5d instr  fp64-pipe %4d  %s" % (name, sum(c.values()), n64, f))
    return n64 * scale, sum(c.values()) * scale


print("outer loop %#x..%#x  stage loop %#x..%#x  rhs at %#x" % (outer + stage + (rhs0,)))
a = show("refill+prestep", outer[0], stage[0])
b = show("stage-state (6 variants)", stage[0], rhs0)
c = show("rhs + store", rhs0, stage[1] + 16, 6.0)
d = show("error/controller/events", stage[1] + 16, outer[1] + 16)
print("per attempt (estimate): fp64-pipe instr %.0f, all instr ~%.0f" % (a[0] + b[0] + c[0] + d[0], a[1] + b[1] / 6 * 1 + c[1] + d[1]))
