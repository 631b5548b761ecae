#!/usr/bin/env python
"""Kernel-by-kernel comparison of the SASS of two builds of libb2r.so (cuobjdump -sass; no GPU needed).

    python tools/sass_diff.py old/libb2r.so new/libb2r.so [-v]

Used when code is added without a GPU at hand (experimental options, new template parameters): a kernel that prints
SAME has the identical instruction sequence as the build that was measured; `same but for constant-bank offsets /
registers` means only kernel-parameter offsets or register numbers moved (e.g. a field added to an argument struct);
anything else lists the differing opcodes.  Template instantiations are matched by their leading arguments, so
`raster_shade_kernel<a,b,c>` pairs with `raster_shade_kernel<a,b,c,false>`.
A checkout of an older commit is built with:  git archive <commit> custom_gpgpu_rasterizer_b200/csrc include | tar -x -C /tmp/old
&& make -C /tmp/old/custom_gpgpu_rasterizer_b200/csrc libb2r.so"""
import difflib
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, name, lines = {}, None, []
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            if name:
                res[name] = lines
            name, lines = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            lines.append(re.sub(r"/\*[0-9a-fx]+\*/", "", l).strip())
    if name:
        res[name] = lines
    names = subprocess.run(["cu++filt"] + list(res), capture_output=True, text=True, check=True).stdout.splitlines()
    return {d: res[k] for d, k in zip(names, res)}


def norm(ls):
    out = []
    for l in ls:
        l = re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0][X]", l)
        l = re.sub(r"\bU?R\d+\b", "R", l)
        l = re.sub(r"\bU?P\d\b", "P", l)
        out.append(re.sub(r"0x[0-9a-f]+", "IMM", l))
    return out


def opcode(l):
    t = l.split()
    return (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0]) if t else ""


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    verbose = "-v" in sys.argv[3:]
    used = set()
    for ka in sorted(a):
        kb = ka if ka in b else None
        if kb is None and ">(" in ka:       # a template that gained trailing parameters
            head = ka.split(">(")[0]
            cand = [k for k in b if k.startswith(head + ", ") and k.split(">(")[0][len(head):].replace(", (bool)0", "") == ""]
            kb = cand[0] if cand else None
        if kb is None:                       # same function name, other signature
            cand = [k for k in b if k.split("(")[0] == ka.split("(")[0] and k not in a]
            kb = cand[0] if len(cand) == 1 else None
        if kb is None:
            print(f"GONE    {len(a[ka]):5d}        {ka[:110]}")
            continue
        used.add(kb)
        if a[ka] == b[kb]:
            verdict = "SAME"
        elif norm(a[ka]) == norm(b[kb]):
            verdict = "same but for constant-bank offsets / registers"
        else:
            oa, ob = [opcode(x) for x in a[ka]], [opcode(x) for x in b[kb]]
            sm = difflib.SequenceMatcher(None, oa, ob, autojunk=False)
            verdict = f"DIFFERS (opcode similarity {sm.ratio():.4f})"
            if verbose:
                for tag, i1, i2, j1, j2 in sm.get_opcodes():
                    if tag != "equal":
                        verdict += f"\n          {tag}: {[x.split(';')[0] for x in a[ka][i1:i2][:4]]} -> {[x.split(';')[0] for x in b[kb][j1:j2][:4]]}"
        print(f"{verdict.splitlines()[0][:56]:56s} {len(a[ka]):5d} {len(b[kb]):5d}  {ka[:90]}")
        for extra in verdict.splitlines()[1:]:
            print(extra[:300])
    for kb in sorted(set(b) - used):
        print(f"NEW           {len(b[kb]):5d}  {kb[:110]}")


if __name__ == "__main__":
    main()
