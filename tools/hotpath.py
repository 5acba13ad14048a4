#!/usr/bin/env python3
"""Developer tool: the executed hot path of trace_kernel from an ncu source capture merged with the cuobjdump SASS of
the same binary (which carries the .reuse flags ncu drops).  Prints instructions / model cycles per warp pass
(one pass of the outer loop = one step attempt of the warp's lanes) and, with --list, every instruction executed
at least --min times per pass.
usage: hotpath.py <lib.so | kernel.sass> <source_sass.csv> [--list] [--min 0.3] [--nofp64] [--kernel trace_kernelILi1ELi0]"""
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import importlib.util
spec = importlib.util.spec_from_file_location("rm", os.path.join(os.path.dirname(os.path.abspath(__file__)), "regread_model.py"))


def load_model():
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "regread_model.py")).read()
    src = src.rsplit("\nmain()", 1)[0]
    ns = {}
    exec(compile(src, "regread_model", "exec"), ns)
    return ns


M = load_model()
arg = sys.argv[1]
kern = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "trace_kernelILi1ELi0"
if arg.endswith(".so"):
    txt = subprocess.run(["cuobjdump", "-sass", arg], capture_output=True, text=True).stdout
    ins, on = [], False
    for l in txt.split("\n"):
        if "Function :" in l:
            on = kern in l
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
else:
    ins = M["parse_sass"](arg)
rows = list(csv.reader(open(sys.argv[2])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
prof = [r for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
assert len(prof) == len(ins), "SASS and capture are of different binaries: %d vs %d instructions" % (len(ins), len(prof))
execd = [int(r[col["Instructions Executed"]] or 0) for r in prof]
smp = [int(r[col["# Samples"]] or 0) for r in prof]
cs = M["cost_stream"](ins)
back = []
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA (0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        back.append((int(m.group(1), 16), a, execd[k]))
outer = max(back, key=lambda b: (b[1] - b[0]) * (b[2] > 0))
passes = float(outer[2])
minx = float(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0.3
tot_i = tot_c = 0.0
cls = collections.defaultdict(lambda: [0.0, 0.0])
for (a, t, base, cyc, words, eff), e in zip(cs, execd):
    tot_i += e
    tot_c += e * cyc
    k = "fp64" if base in M["FP64"] else base
    cls[k][0] += e
    cls[k][1] += e * cyc
print("outer loop %#x..%#x, %.4g warp passes; %.1f instr / pass, %.0f model cycles / pass; %d samples"
      % (outer[0], outer[1], passes, tot_i / passes, tot_c / passes, sum(smp)))
print("per class (instr/pass, model cycles/pass):")
for k, (e, c) in sorted(cls.items(), key=lambda kv: -kv[1][1])[:30]:
    print("  %-10s %8.1f %8.1f" % (k, e / passes, c / passes))
if "--list" in sys.argv:
    for (a, t, base, cyc, words, eff), e, s in zip(cs, execd, smp):
        if e / passes < minx:
            continue
        if "--nofp64" in sys.argv and base in ("DFMA", "DMUL", "DADD"):
            continue
        print("%05x %-64s x%5.2f cyc %.1f smp %d" % (a, t[:64], e / passes, cyc, s))
