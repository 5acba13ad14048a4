#!/usr/bin/env python3
"""Developer tool: static model cost (tools/regread_model.py) of one whole kernel of the library, every instruction
counted once -- meant for straight-line kernels such as rhs_kernel.  usage: kernel_cost.py <lib.so> <kernel substring>"""
import os, re, subprocess, sys, collections
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "regread_model.py")).read().rsplit("\nmain()", 1)[0]
M = {}
exec(compile(src, "regread_model", "exec"), M)
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
ins, on = [], False
for l in txt.split("\n"):
    if "Function :" in l:
        on = sys.argv[2] in l
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
cs = M["cost_stream"](ins)
c = collections.Counter(); n = collections.Counter()
three = hit = 0
for a, t, base, cyc, words, eff in cs:
    c[base] += cyc; n[base] += 1
    if base == "DFMA" and words == 6:
        three += 1
        if eff < 6: hit += 1
tot = sum(c.values())
print("%d instr, %.1f model cycles; FP64 %d instr %.1f cycles; 3-reg DFMA %d (reuse hits %d)" % (
    len(ins), tot, sum(n[k] for k in M["FP64"]), sum(c[k] for k in M["FP64"]), three, hit))
