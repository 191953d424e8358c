#!/usr/bin/env python
"""Summarise an .ncu-rep (first kernel): headline metrics, pipe utilisation, stall mix, top opcodes.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))


def g(k):
    return m.get(k, "?")


print("kernel:", g("Kernel Name")[:80], " grid", g("launch__grid_size"), "block", g("launch__block_size"),
      "regs", g("launch__registers_per_thread"), "dyn smem", g("launch__shared_mem_per_block_dynamic"))
print("duration", g("gpu__time_duration.sum"), hdr and rows[1][hdr.index("gpu__time_duration.sum")],
      "| dram read", g("dram__bytes_read.sum"), rows[1][hdr.index("dram__bytes_read.sum")], "write", g("dram__bytes_write.sum"),
      rows[1][hdr.index("dram__bytes_write.sum")], "| dram %", g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
      "l2 %", g("lts__throughput.avg.pct_of_peak_sustained_elapsed"), "l1 %", g("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"))
print("inst", g("smsp__inst_executed.sum"), "ipc", g("sm__inst_executed.avg.per_cycle_active"), "issue active %",
      g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps active %", g("sm__warps_active.avg.pct_of_peak_sustained_active"))
print("pipes %:", {k.split("pipe_")[1].split(".")[0]: round(float(v), 1) for k, v in m.items()
                    if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") and float(v) > 0.5})
print("stalls/issue:", {k.split("issue_stalled_")[1].split("_per_")[0]: round(float(v), 2) for k, v in m.items()
                        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(v) > 0.05})
print("smem wavefronts", g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), "bank conflicts ld", g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
agg = defaultdict(lambda: [0, 0])
for r in data:
    s = r[ix["Source"]].split()
    op = (s[1] if s and s[0].startswith("@") else (s[0] if s else "")).split(".")[0]
    agg[op][0] += int(r[ix["# Samples"]] or 0)
    agg[op][1] += int(r[ix["Instructions Executed"]] or 0)
print("samples by opcode:", [(op, f"{100 * s / tot:.1f}%", e) for op, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:10]])
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tops = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in tops:
    st = {k[6:]: int(r[ix[k]] or 0) for k in stall_cols if int(r[ix[k]] or 0) > 0}
    print(" ", r[ix["Address"]][-5:], r[ix["Source"]][:56].strip().ljust(56), r[ix["# Samples"]].rjust(6), dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
