#!/usr/bin/env python
"""SURVEY section 8(f) rank 1, measured: the product's hash-based ObjParser against the reference's O(n^2) parser
(oracle/_ref/ref_render parse / parsev = the two ReadFromObjFile overloads of the reference's ObjParser.cpp, CPU only - runs
without a GPU; the pointer overload overruns its buffer on meshes whose corners rarely share vertices, ObjParser.cpp:212, so the
vector overload is timed except for the position-only bunny), same OBJ text, and whether
the two vertex / index buffers are identical.   python tools/objparser_bench.py [grid sizes ...]   (default 40 80 200 400)"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases  # noqa: E402
from custom_gpgpu_rasterizer_b200.renderer import parse_obj  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
RES = os.path.join(ROOT, "oracle", "_ref", "resources")


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [40, 80, 200, 400]
    with tempfile.TemporaryDirectory() as td:
        cases = [(os.path.join(RES, f), inv, vec) for f, inv, vec in (("vehicle.obj", True, True), ("tuktuk.obj", True, True), ("lowpoly_bunny.obj", False, False)) if os.path.exists(os.path.join(RES, f))]
        for n in sizes:
            p = os.path.join(td, f"grid{n}.obj")
            with open(p, "w") as f:
                f.write(golden_cases._obj_grid(n, True, True, seed=5))
            cases.append((p, True, True))
        print("| mesh | vertices | triangles | this parser | reference parser (incl. ~20 ms process start) | ratio | buffers identical |\n|---|---|---|---|---|---|---|")
        for p, inv, vec in cases:
            t = time.perf_counter()
            v, i, _ = parse_obj(p, invert_y=inv, pointer_flavour=not vec)
            ours = time.perf_counter() - t
            out = os.path.join(td, "o.bin")
            t = time.perf_counter()
            subprocess.run([REF, "parsev" if vec else "parse", p, str(int(inv)), out], stdout=subprocess.DEVNULL, check=True)
            ref = time.perf_counter() - t
            raw = np.fromfile(out, dtype=np.uint8)
            hdr = raw[:16].view(np.uint32)
            nv, ni = int(hdr[1]), int(hdr[2])
            rv = raw[16:16 + nv * 56].view(np.uint32).reshape(nv, 14)
            ri = raw[16 + nv * 56:16 + nv * 56 + ni * 4].view(np.uint32)
            same = nv == v.shape[0] and np.array_equal(ri, i) and np.array_equal(rv, np.ascontiguousarray(v, np.float32).view(np.uint32).reshape(nv, 14))
            print(f"| {os.path.basename(p)} | {nv} | {ni // 3} | {ours * 1e3:.1f} ms | {ref * 1e3:.1f} ms | {ref / ours:.1f}x | {same} |")


if __name__ == "__main__":
    main()
