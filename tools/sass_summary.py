#!/usr/bin/env python
"""Opcode histogram of every kernel in libb2r.so (cuobjdump -sass), as markdown: which memory / atomic / tensor / TMA / PDL
instructions the shipped cubin really contains.  No GPU needed.
usage: tools/sass_summary.py [libb2r.so] > profiles/rNN_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "custom_gpgpu_rasterizer_b200/csrc/libb2r.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
names = {}
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
dem = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
GROUPS = [("global loads", r"^LDG(?!STS)"), ("async copies global->shared (LDGSTS = cp.async)", r"^LDGSTS"), ("global stores", r"^STG"), ("shared ld/st", r"^(LDS|STS)"), ("local (spill) ld/st", r"^(LDL|STL)"),
          ("global atomics / reductions", r"^(ATOMG|ATOM\b|ATOM\.|RED)"), ("shared atomics", r"^ATOMS"), ("warp collectives", r"^(SHFL|VOTE|MATCH|REDUX|CREDUX)"),
          ("barriers", r"^BAR"), ("prefetch (CCTL / PREFETCH)", r"^(CCTL|PREFETCH)"), ("PDL (ACQBULK / griddepcontrol: DEPBAR-less)", r"^(ACQBULK|PREEXIT)"),
          ("TMA (UBLKCP / UTMALDG / UTMASTG)", r"^(UBLKCP|UTMA)"), ("tensor core (HMMA / IMMA / UTCMMA / tcgen05)", r"^(HMMA|IMMA|DMMA|QMMA|UTC)"),
          ("MUFU", r"^MUFU"), ("FP32 FMA/MUL/ADD", r"^(FFMA|FMUL|FADD)\b"), ("integer IMAD/IADD/LOP", r"^(IMAD|IADD3|VIADD|LOP3)")]
print(f"# SASS summary of `{lib}` ({', '.join(arch)})\n")
print("One row per kernel: static instruction counts by class (cuobjdump -sass).  No TMA and no tensor-core instruction appears in any kernel (the setup kernel's copy pipeline is cp.async = LDGSTS: its sources are gathers, not tiles a tensor map could describe) -")
print("the pipeline is gather / scatter / per-pixel arithmetic (DESIGN.md section 4); programmatic dependent launch shows as ACQBULK / PREEXIT-style control in the prologues.\n")
print("| kernel | instructions | " + " | ".join(g for g, _ in GROUPS) + " |")
print("|---|---|" + "---|" * len(GROUPS))
tot = collections.Counter()
for (mangled, hist), name in zip(kernels.items(), dem):
    short = re.sub(r"\(b2r::.*|\((?:unsigned|const|float|int|uint).*", "", name).replace("void ", "").replace("b2r::", "").replace("(bool)", "")
    row = []
    for g, pat in GROUPS:
        n = sum(c for op, c in hist.items() if re.match(pat, op))
        tot[g] += n
        row.append(str(n))
    print(f"| {short} | {sum(hist.values())} | " + " | ".join(row) + " |")
print("\nTotals: " + ", ".join(f"{g}: {n}" for g, n in tot.items()))
allops = collections.Counter()
for h in kernels.values():
    allops.update(h)
print("\nAll opcodes (static count over all kernels): " + ", ".join(f"{op} {n}" for op, n in allops.most_common()))
