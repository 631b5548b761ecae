#!/usr/bin/env python
"""Per CUDA source line: share of executed warp instructions of one kernel.  usage: tools/ncu_instr.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; L = []
def num(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or not r[0].isdigit() or len(r) < len(hdr) - 2: continue
    off = len(r) - len(hdr)       # a source text containing commas inside quotes is handled by csv; guard anyway
    L.append((cur, int(r[0]), r[1], num(r[ii + off]), num(r[ti + off]), num(r[si + off])))
tot = sum(x[3] for x in L) or 1
print("total warp instructions %.0f, thread instructions %.0f, avg active threads %.1f" % (tot, sum(x[4] for x in L), sum(x[4] for x in L) / tot))
for f, ln, src, i, t, s in sorted(L, key=lambda x: -x[3])[:n]:
    print(f"{i / tot * 100:5.1f}%  thr/warp {t / max(i, 1):5.1f}  {f}:{ln:<4} {src.strip()[:100]}")
