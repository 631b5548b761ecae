#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ from the REFERENCE ITSELF.

  python tools/make_golden.py parser     (here: runs the reference ObjParser, CPU only, needs /root/reference
                                          or the prebuilt oracle/_ref/ref_render)
  python tools/make_golden.py frames     (GPU box: runs the unmodified reference CUDA renderer
                                          oracle/_ref/ref_render on small scenes; write into gpurun_out/golden/)

  python tools/make_golden.py erenderer  (here, CPU only: runs the reference's own CPU renderer, ERenderer.cpp built by
                                          oracle/build_ref_erenderer.sh into oracle/_ref/libere_ref.so, on the frame cases and on
                                          BASELINE config 1 = the bundled textured vehicle at 640x480)

The scenes / OBJ texts are regenerated deterministically by tests/golden_cases.py, so only the
reference's OUTPUTS are stored.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_render")
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_parse(obj_path, invert_y, vector_overload=False):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "o.bin")
        subprocess.run([REF, "parsev" if vector_overload else "parse", obj_path, str(int(invert_y)), out], check=True, stdout=subprocess.DEVNULL)
        raw = np.fromfile(out, dtype=np.uint8)
    hdr = raw[:16].view(np.uint32)
    assert hdr[0] == 0x50523242
    nv, ni, vt = int(hdr[1]), int(hdr[2]), int(hdr[3])
    v = raw[16:16 + nv * 56].view(np.float32).reshape(nv, 14).copy()
    i = raw[16 + nv * 56:16 + nv * 56 + ni * 4].view(np.uint32).copy()
    return v, i, vt


def parser():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (text, invert) in golden_cases.obj_cases().items():
            p = os.path.join(td, name + ".obj")
            with open(p, "w") as f:
                f.write(text)
            v, i, vt = ref_parse(p, invert)
            out[name + "_v"], out[name + "_i"], out[name + "_t"] = v, i, np.array([vt])
            v, i, vt = ref_parse(p, invert, True)
            out[name + "_vec_v"], out[name + "_vec_i"], out[name + "_vec_t"] = v, i, np.array([vt])
    # the reference's own bundled meshes (inputs stay in /root/reference; only the parser output is stored)
    res = "/root/reference/CustomCUDARasterizer/Resources"
    if os.path.isdir(res):
        for name, invert, vec in (("vehicle", True, True), ("tuktuk", True, True), ("lowpoly_bunny", False, False)):
            v, i, vt = ref_parse(os.path.join(res, name + ".obj"), invert, vec)
            out["res_" + name + "_v"], out["res_" + name + "_i"], out["res_" + name + "_t"] = v, i, np.array([vt])
    np.savez_compressed(os.path.join(GOLD, "objparser_reference.npz"), **out)
    print("wrote", os.path.join(GOLD, "objparser_reference.npz"), {k: v.shape for k, v in out.items()})


def frames():
    import parity
    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    for name, make in golden_cases.frame_cases().items():
        sc = make()
        ref = parity.run_ref(sc, "stages")
        rows = []
        for mi, m in enumerate(ref["meshes"]):
            p = m["bin_pairs"]
            if p.size:
                rows.append(np.concatenate([p[:, :1], np.full((p.shape[0], 1), mi, np.uint32), p[:, 1:]], 1))
        bins = parity.sort_rows(np.concatenate(rows, 0)) if rows else np.zeros((0, 5), np.uint32)
        np.savez_compressed(os.path.join(dst, f"frame_{name}.npz"),
                            framebuffer=ref["framebuffer"], depth=ref["depth"], z=ref["z"],
                            view_proj=ref["view_proj"], cam_pos=ref["cam_pos"], cam_fwd=ref["cam_fwd"],
                            visible_tris=np.array([ref["visible_tris"]]), total_tris=np.array([ref["total_tris"]]),
                            bins=bins,
                            overtex0=ref["meshes"][0]["overtex"],
                            visible0=parity.sort_rows(ref["meshes"][0]["visible_triples"]))
        print("wrote", name, ref["visible_tris"], bins.shape)


def erenderer():
    """tests/golden/erenderer_reference.npz: SHA-256 of back buffer and depth buffer (float bits) of every frame case as
    rendered by the reference's Elite::Renderer, the full surfaces of three of them (diagnosable failures), and the
    same for BASELINE config 1 (bundled vehicle, four maps, 640x480, camera {0,5,65}) when the assets are present."""
    import hashlib
    import asset_scenes
    import oracle_lib
    assert oracle_lib.have_erenderer_reference(), "run oracle/build_ref_erenderer.sh first"
    cases = dict(golden_cases.frame_cases())
    if asset_scenes.have_assets():
        cases["config1_vehicle_640x480"] = lambda: asset_scenes.vehicle_scene(640, 480, sample=0)
        cases["config1_vehicle_640x480_linear"] = lambda: asset_scenes.vehicle_scene(640, 480, sample=1)
    out = {}
    for name, make in cases.items():
        r = oracle_lib.erenderer_render(make(), reference=True)
        fb, depth = r["framebuffer"], r["depth"].view(np.uint32)
        out[name + "_sha"] = np.array([hashlib.sha256(fb.tobytes()).hexdigest(), hashlib.sha256(depth.tobytes()).hexdigest()])
        out[name + "_drawn"] = np.array([int((~np.isnan(r["depth"])).sum())])
        if name in ("soup_textured_linear", "quads_linear", "spheres"):
            out[name + "_fb"], out[name + "_depth"] = fb, depth
        print(name, out[name + "_drawn"][0], "pixels drawn", f"{r['seconds'] * 1e3:.1f} ms")
    np.savez_compressed(os.path.join(GOLD, "erenderer_reference.npz"), **out)
    print("wrote", os.path.join(GOLD, "erenderer_reference.npz"))


if __name__ == "__main__":
    {"parser": parser, "frames": frames, "erenderer": erenderer}[sys.argv[1]]()
