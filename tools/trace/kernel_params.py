#!/usr/bin/env python
"""Kernel-parameter layouts (offset, size per ordinal) of every kernel in a library, read from the cubin's .nv.info
sections (`cuobjdump -elf`).  Written as the table tools/trace/fake_cudart.cpp loads to dump launch arguments:
one line per kernel:  <mangled name> <n> <offset0> <size0> ... ."""
import re
import subprocess
import sys


def kernel_params(lib):
    out = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True, check=True).stdout
    params, name = {}, None
    for line in out.splitlines():
        m = re.match(r"^\.nv\.info\.(\S+)", line)
        if m:
            name = m.group(1)
            params[name] = {}
            continue
        if line.startswith(".") or line.startswith("Fatbin") or not line.strip():
            if line.startswith(".") and not line.startswith(".nv.info."):
                name = None
            continue
        m = re.search(r"Ordinal : (0x[0-9a-f]+)\s+Offset\s+: (0x[0-9a-f]+)\s+Size\s+: (0x[0-9a-f]+)", line)
        if m and name:
            params[name][int(m.group(1), 16)] = (int(m.group(2), 16), int(m.group(3), 16))
    return {k: [v[i] for i in sorted(v)] for k, v in params.items() if v}


if __name__ == "__main__":
    table = kernel_params(sys.argv[1])
    with open(sys.argv[2], "w") as f:
        for k, v in sorted(table.items()):
            f.write(f"{k} {len(v)} " + " ".join(f"{o} {s}" for o, s in v) + "\n")
    print(f"{len(table)} kernels -> {sys.argv[2]}")
