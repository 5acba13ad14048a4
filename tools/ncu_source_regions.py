#!/usr/bin/env python3
"""Developer tool: summarise the SASS source page of an ncu capture (`ncu -i X.ncu-rep --page source --csv
--print-source sass`) of trace_kernel: executed warp instructions and stall samples per instruction class and
per region of the hot loop, per step attempt.  usage: ncu_source_regions.py source_sass.csv [warp_attempts]"""
import csv
import collections
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
ins = []
base = None
for r in rows[2:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    ins.append(dict(addr=a - base, text=r[col["Source"]].strip(), samples=int(r[col["# Samples"]] or 0),
                    execd=int(r[col["Instructions Executed"]] or 0),
                    thr=int(r[col["Thread Instructions Executed"]] or 0),
                    stalls={k[6:]: int(r[i] or 0) for k, i in col.items() if k.startswith("stall_") and "Not Issued" not in k}))
tot_s = sum(i["samples"] for i in ins)
tot_e = sum(i["execd"] for i in ins)


def op(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]


# hot loop structure from back edges
back = []
for i in ins:
    m = re.search(r"BRA (0x[0-9a-f]+)", i["text"])
    if m:
        tgt = int(m.group(1), 16) - (base if int(m.group(1), 16) >= base else 0)
        if tgt < i["addr"]:
            back.append((tgt, i["addr"]))
outer = max((b for b in back if b[1] - b[0] > 0x3000 and b[1] < 0x9000), key=lambda b: b[1] - b[0])
stage = max((b for b in back if outer[0] < b[0] and b[1] < outer[1]), key=lambda b: b[1] - b[0])
tg = collections.Counter()
for i in ins:
    m = re.search(r"^(?:@!?U?P\d+\s+)?BRA (0x[0-9a-f]+)", i["text"])
    if m and stage[0] <= i["addr"] < stage[1]:
        tg[int(m.group(1), 16) - (base if int(m.group(1), 16) >= base else 0)] += 1
rhs0 = tg.most_common(1)[0][0]
# number of warp-level step attempts = executions of the stage-loop back edge / 6
wa = next(i["execd"] for i in ins if i["addr"] == stage[1]) / 6.0
if len(sys.argv) > 2 and not sys.argv[2].startswith("--"):
    wa = float(sys.argv[2])
print("instructions %d, samples %d, executed warp-instr %.4g, warp-level step attempts %.4g" % (len(ins), tot_s, tot_e, wa))
print("outer %#x..%#x stage loop %#x..%#x rhs %#x" % (outer + stage + (rhs0,)))
regions = [("before loop", 0, outer[0]), ("refill+prestep", outer[0], stage[0]), ("stage-state", stage[0], rhs0),
           ("rhs+store(+init)", rhs0, stage[1] + 16), ("error/ctl/event", stage[1] + 16, outer[1] + 16),
           ("out of line", outer[1] + 16, 1 << 30)]
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
print("%-18s %8s %7s %9s %9s %9s  top stalls" % ("region", "samples", "share", "inst/att", "fp64/att", "smp/inst"))
for name, lo, hi in regions:
    sel = [i for i in ins if lo <= i["addr"] < hi]
    s = sum(i["samples"] for i in sel)
    e = sum(i["execd"] for i in sel)
    f = sum(i["execd"] for i in sel if op(i["text"]) in FP64)
    st = collections.Counter()
    for i in sel:
        st.update(i["stalls"])
    top = ", ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in st.most_common(5))
    print("%-18s %8d %6.1f%% %9.1f %9.1f %9.3f  %s" % (name, s, 100.0 * s / tot_s, e / wa, f / wa, s / max(e, 1) * 1e3, top))
print("\nper opcode (executed per warp-attempt, samples share):")
byop = collections.defaultdict(lambda: [0, 0])
for i in ins:
    byop[op(i["text"])][0] += i["execd"]
    byop[op(i["text"])][1] += i["samples"]
for k, (e, s) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:28]:
    print("  %-10s %8.1f /attempt  %5.1f%% of samples" % (k, e / wa, 100.0 * s / tot_s))
# three-register DFMAs (all distinct registers, no constant/immediate operand) and .reuse flags
n3 = n3r = nd = 0
for i in ins:
    if op(i["text"]) != "DFMA":
        continue
    nd += i["execd"]
    ops_ = i["text"].split(None, 1)[1] if " " in i["text"] else ""
    ops_ = re.sub(r"^DFMA\s+", "", re.sub(r"^@!?U?P\d+\s+", "", i["text"]))
    parts = [p.strip() for p in ops_.split(",")]
    srcs = parts[1:4]
    regs = [re.sub(r"[-|]|\.reuse", "", s) for s in srcs if re.match(r"^-?\|?R\d+", s)]
    if len(regs) == 3 and len(set(regs)) == 3:
        n3 += i["execd"]
        if any(".reuse" in s for s in srcs):
            n3r += i["execd"]
print("\nDFMA per attempt %.1f; with three distinct register sources %.1f (of which carrying a .reuse flag %.1f)" % (nd / wa, n3 / wa, n3r / wa))
if "--top" in sys.argv:
    print("\nhottest instructions:")
    for i in sorted(ins, key=lambda i: -i["samples"])[:60]:
        st = ", ".join("%s %d" % kv for kv in collections.Counter(i["stalls"]).most_common(3))
        print("  %#06x %-52s smp %6d exec/att %6.2f  %s" % (i["addr"], i["text"][:52], i["samples"], i["execd"] / wa, st))
if "--dfma3" in sys.argv:
    print("\nthree-register DFMAs executed at least 0.5x per attempt, in address order (with neighbours' opcodes):")
    for k, i in enumerate(ins):
        if op(i["text"]) != "DFMA" or i["execd"] / wa < 0.5:
            continue
        ops_ = re.sub(r"^DFMA\s+", "", re.sub(r"^@!?U?P\d+\s+", "", i["text"]))
        srcs = [p.strip() for p in ops_.split(",")][1:4]
        regs = [re.sub(r"[-|]|\.reuse", "", s) for s in srcs if re.match(r"^-?\|?R\d+", s)]
        if len(regs) == 3 and len(set(regs)) == 3:
            print("  %#06x %-46s x%.2f smp %d" % (i["addr"], i["text"][:46], i["execd"] / wa, i["samples"]))
if "--list" in sys.argv:
    lo, hi = [int(v, 16) for v in sys.argv[sys.argv.index("--list") + 1].split(":")]
    for i in ins:
        if lo <= i["addr"] < hi:
            print("%#06x %-60s x%.3f smp %d" % (i["addr"], i["text"][:60], i["execd"] / wa, i["samples"]))
