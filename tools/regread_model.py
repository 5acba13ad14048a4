#!/usr/bin/env python3
"""Developer tool: static cost model of the trace kernel's hot loop on B200.

Measured (tests/microbench_fp64.py, profiles/r02b_fp64_modes.log): a scheduler's register file delivers about two
32-bit words per cycle; a DFMA with three distinct register sources takes 2.88 cycles instead of 2, and every integer
instruction interleaved with full-rate DFMAs costs 1-1.5 further cycles.  Model: each executed warp-instruction costs
max(dispatch cycles of its pipe, register source words / 2); operands served by the operand-reuse cache, constants,
immediates, uniform registers, RZ and predicates are free.

usage: regread_model.py <cuobjdump-style .sass of the kernel> [ncu source csv of the SAME binary]
Without the csv the hot-loop trip counts are assumed (stage loop x6, everything else x1 when on the common path is
unknown -> printed per region statically)."""
import collections
import csv
import re
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def parse_sass(path):
    ins = []
    for l in open(path):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]


NO_DEST = ("ST", "STS", "STL", "STG", "BRA", "BRX", "EXIT", "BSYNC", "BSSY", "RET", "CALL", "RED", "REDG", "ATOM", "WARPSYNC",
           "NOP", "BAR", "MEMBAR", "ERRBAR", "ENDCOLLECTIVE", "YIELD", "BREAK", "BMOV")
TWO_DEST = ("ISETP", "DSETP", "FSETP", "PLOP3", "UISETP", "IADD3", "VOTE", "R2P", "SHFL", "UIADD3", "LEA", "IMAD.WIDE")


def sources(t):
    """-> (opcode base, list of (slot, reg, words, reuse_flag))"""
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    parts = t.split(None, 1)
    full = parts[0]
    base = full.split(".")[0]
    if len(parts) == 1:
        return base, full, []
    ops = [o.strip() for o in parts[1].split(",")]
    if base in NO_DEST:
        src = ops
    else:
        nd = 1
        # predicate destinations
        while nd < len(ops) and re.match(r"^!?U?P(T|\d+)$", ops[nd]) and base in ("ISETP", "DSETP", "FSETP", "PLOP3", "UISETP", "VOTE"):
            nd += 1
            if nd >= 2 and base != "PLOP3":
                break
        if base in ("IADD3", "UIADD3", "LEA", "VIMNMX", "VIADDMNMX"):   # optional predicate carry-outs
            while nd < len(ops) and re.match(r"^!?U?P(T|\d+)$", ops[nd]):
                nd += 1
        src = ops[nd:]
    wide = 2 if (base in FP64 or base in ("DMNMX",)) else 1
    out = []
    for k, o in enumerate(src):
        reuse = ".reuse" in o
        o2 = o.replace(".reuse", "")
        m = re.search(r"(?<![UP\w])R(\d+)(\.64)?", o2)
        if not m or re.match(r"^-?\|?RZ", o2.strip("[]")):
            continue
        w = wide
        if "[" in o2:      # address operand
            w = 2 if m.group(2) else 1
        elif base in ("STS", "STL", "ST", "STG"):
            w = 4 if ".128" in full else (2 if ".64" in full else 1)
        elif base == "IMAD" and ".WIDE" in full and k == 2:
            w = 2
        elif base == "MUFU":
            w = 1
        elif base in ("F2F", "I2F", "F2I", "FRND") and ".F64" in full:
            w = 2 if full.endswith("F64") or ".F64." in full.split(".", 1)[1][4:] else 1
        out.append((k, int(m.group(1)), w, reuse))
    return base, full, out


def dispatch_cycles(base, full):
    if base in FP64:
        return 2.0
    if base == "MUFU":
        return 1.0
    return 1.0


def cost_stream(ins):
    """-> per instruction (cycles, words, words_after_reuse)"""
    cache = {}
    res = []
    for a, t in ins:
        base, full, src = sources(t)
        words = 0
        eff = 0
        used = set()
        seen = set()
        for k, r, w, reuse in src:
            used.add(k)
            if r not in seen:      # the same register in two slots is read once (a*a+c runs at the full rate)
                words += w
                if cache.get(k) != r:
                    eff += w
                seen.add(r)
            if reuse:
                cache[k] = r
            else:
                cache.pop(k, None)
        res.append((a, t, base, max(dispatch_cycles(base, full), eff / 2.0), words, eff))
    return res


def main():
    ins = parse_sass(sys.argv[1])
    cs = cost_stream(ins)
    execd = None
    if len(sys.argv) > 2 and not sys.argv[2].startswith("--"):
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        ex = [int(r[col["Instructions Executed"]] or 0) for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
        assert len(ex) == len(ins), (len(ex), len(ins))
        execd = ex
    # loop structure
    back = []
    for a, t in ins:
        m = re.search(r"BRA (0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            back.append((int(m.group(1), 16), a))
    outer = max((b for b in back if b[1] - b[0] > 0x3000 and b[1] < 0x9000), key=lambda b: b[1] - b[0])
    stage = max((b for b in back if outer[0] < b[0] and b[1] < outer[1]), key=lambda b: b[1] - b[0])
    if execd:
        wa = execd[[a for a, _ in ins].index(stage[1])] / 6.0
        tot = collections.Counter()
        byop = collections.defaultdict(float)
        n = collections.Counter()
        for (a, t, base, cyc, words, eff), e in zip(cs, execd):
            reg = ("pre" if a < outer[0] else "refill+prestep" if a < stage[0] else "stage loop" if a <= stage[1]
                   else "error/ctl/event" if a <= outer[1] else "out of line")
            tot[reg] += e * cyc
            n[reg] += e
            byop[base] += e * cyc
        print("warp-level attempts %.4g" % wa)
        s = 0
        for k in ("pre", "refill+prestep", "stage loop", "error/ctl/event", "out of line"):
            print("  %-16s %8.1f instr/attempt  %8.1f model cycles/attempt" % (k, n[k] / wa, tot[k] / wa))
            s += tot[k] / wa
        print("  total model cycles per warp-attempt: %.0f" % s)
        print("  by opcode:", ", ".join("%s %.0f" % (k, v / wa) for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:18]))
    else:
        tot = collections.Counter()
        for a, t, base, cyc, words, eff in cs:
            reg = ("pre" if a < outer[0] else "refill+prestep" if a < stage[0] else "stage loop" if a <= stage[1]
                   else "error/ctl/event" if a <= outer[1] else "out of line")
            tot[reg] += cyc
        print("static model cycles by region (every instruction once):", dict(tot))


main()
