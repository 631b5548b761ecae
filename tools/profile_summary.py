#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep, `--set full`) into a small markdown table that is committed
under profiles/.  usage: tools/profile_summary.py report.ncu-rep [title] > profiles/xyz.md"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, body = rows[0], rows[1], rows[2:]


def col(name):
    return h.index(name) if name in h else None


def val(r, name, fmt="{:.1f}"):
    i = col(name)
    if i is None or r[i] == "":
        return "-"
    try:
        return fmt.format(float(r[i].replace(",", "")))
    except ValueError:
        return r[i]


stall_cols = [(i, n.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, n in enumerate(h)
              if "pcsamp_warps_issue_stalled" in n and not n.endswith("_not_issued")]
print(f"# {title}\n")
print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch, cold-cache, serialised)\n")
print("| kernel | time us | DRAM rd MB | DRAM wr MB | DRAM % | SM % | occupancy % | regs | thr/inst | L1 hit % | L2 hit % | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in body:
    name = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("b2r::", "")
    st = [(float(r[i]) if r[i] else 0.0, n) for i, n in stall_cols]
    tot = sum(v for v, _ in st) or 1.0
    top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:4])
    print("| " + " | ".join([
        name, val(r, "gpu__time_duration.sum", "{:.1f}"),
        val(r, "dram__bytes_read.sum", "{:.1f}"), val(r, "dram__bytes_write.sum", "{:.1f}"),
        val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), val(r, "launch__registers_per_thread", "{:.0f}"),
        val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), val(r, "l1tex__t_sector_hit_rate.pct"),
        val(r, "lts__t_sector_hit_rate.pct"), top]) + " |")
print()
print("units: time " + (units[col("gpu__time_duration.sum")] if col("gpu__time_duration.sum") is not None else "?") +
      ", DRAM " + (units[col("dram__bytes_read.sum")] if col("dram__bytes_read.sum") is not None else "?"))
