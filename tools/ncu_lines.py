#!/usr/bin/env python
"""Per CUDA source line: warp-stall samples, executed warp instructions and the dominant stall reasons of one kernel.
usage: tools/ncu_lines.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    # the second "Source" column shadows the first in dict(); take the text by position
    lines.append((cur_file, int(r[0]), r[1], d))
def num(x):
    try: return float(x)
    except Exception: return 0.0
tot_s = sum(num(d["# Samples"]) for *_, d in lines) or 1
tot_i = sum(num(d["Instructions Executed"]) for *_, d in lines) or 1
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
agg = {}
for k in stalls: agg[k] = sum(num(d[k]) for *_, d in lines)
print("stalls:", ", ".join(f"{k[6:]} {v / tot_s * 100:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for f, ln, src, d in sorted(lines, key=lambda t: -num(t[3]["# Samples"]))[:n]:
    top = sorted(((num(d[k]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{num(d['# Samples']) / tot_s * 100:5.1f}% smp {num(d['Instructions Executed']) / tot_i * 100:5.1f}% ins  thr {d['Avg. Threads Executed']:>3}  {f}:{ln:<4} {src.strip()[:90]}   [{', '.join(f'{k} {v:.0f}' for v, k in top if v)}]")
