#!/usr/bin/env python3
"""Developer tool: condense one `ncu --set full` capture of trace_kernel (raw page CSV) plus the bench line of the
same build and workload into the small JSON that bench.py quotes beside its live numbers.
usage: ncu_summary.py <raw.csv> <bench line .json (same build, same workload)> <out.json> [note]"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

raw, bench, out = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))


def f(name):
    return float(m[name].replace(",", ""))


line = json.loads(open(bench).read().strip().splitlines()[-1])
attempts = line["work"]["step_attempts"]


def thread_inst(op):
    """Thread instructions of one FP64 opcode over the launch (the raw page of `--set full` carries them per elapsed
    cycle: .sum.per_cycle_elapsed x smsp__cycles_elapsed.avg)."""
    name = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op
    if name in m:
        return f(name)
    return f(name + ".per_cycle_elapsed") * f("smsp__cycles_elapsed.avg")


dfma, dmul, dadd = thread_inst("dfma"), thread_inst("dmul"), thread_inst("dadd")
pkg = entry.load_package()
unit = {n: u for n, u in zip(hdr, rows[1])}


def mbytes(name):
    v, u = f(name), unit[name]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


s = {
    "kernel": m["Kernel Name"],
    "workload": line["config"]["workload"], "ni": line["config"]["ni"], "nj": line["config"]["nj"],
    "kernel_source_sha16": pkg._lib.kernel_source_sha16(),
    "gpu_time_ms_under_ncu": f("gpu__time_duration.sum") * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit["gpu__time_duration.sum"]],
    "sm_clock_ghz_under_ncu": f("sm__cycles_elapsed.max.per_second"),
    "kernel_ms_bench": line["roofline"]["kernel_ms"],
    "step_attempts": attempts, "rhs_evals": line["work"]["rhs_evals"],
    "thread_inst_dfma": dfma, "thread_inst_dmul": dmul, "thread_inst_dadd": dadd,
    "fp64_thread_inst_per_attempt": (dfma + dmul + dadd) / attempts,
    "fp64_flops_per_attempt": (2 * dfma + dmul + dadd) / attempts,
    "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "warps_eligible_per_cycle": f("smsp__warps_eligible.avg.per_cycle_active"),
    "warp_inst_executed": f("smsp__inst_executed.sum"), "warp_inst_per_warp_attempt": None,
    "warp_execution_efficiency": f("smsp__thread_inst_executed_per_inst_executed.ratio") / 32.0,
    "registers_per_thread": int(f("launch__registers_per_thread")),
    "dram_bytes_read": mbytes("dram__bytes_read.sum"), "dram_bytes_write": mbytes("dram__bytes_write.sum"),
    "source": os.path.basename(raw),
    "note": sys.argv[4] if len(sys.argv) > 4 else "",
}
s["dram_bytes_per_launch"] = s["dram_bytes_read"] + s["dram_bytes_write"]
# executed warp instructions per step attempt of a warp (a warp attempt = 32 x efficiency thread attempts)
s["warp_inst_per_warp_attempt"] = s["warp_inst_executed"] / (attempts / (32.0 * s["warp_execution_efficiency"]))
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))
