#!/usr/bin/env python
"""Print the hottest SASS lines (by warp-stall samples) of one kernel from an .ncu-rep file.
usage: tools/ncu_hot.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
body = [r for r in rows[hi + 1:] if len(r) > si and r[0] != "Address" and r[si].replace(".", "").isdigit()]
tot = sum(float(r[si] or 0) for r in body) or 1
print("total samples", tot, "instructions", len(body))
for r in sorted(body, key=lambda r: -float(r[si] or 0))[:n]:
    print(f"{float(r[si] or 0) / tot * 100:5.1f}%  {r[ie]:>10}  {r[src][:110]}")
