"""Debug aid: where do product and oracle disagree on the hostile-input scene (reference fill rule)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
from custom_gpgpu_rasterizer_b200 import _lib, scenes
from custom_gpgpu_rasterizer_b200.renderer import Renderer
for fill in (0, 1):
    sc = scenes.hostile_scene(640, 480, 600, seed=21)
    r = Renderer(sc.width, sc.height); r.set_option(_lib.OPT_FILL_RULE_TOP_LEFT, fill); r.set_option(_lib.OPT_DEBUG_SURFACES, 1)
    r.load_scene(sc); r.render_scene(sc)
    o = oracle_lib.render(sc, fill_rule=fill, model_holes=False)
    w, d, z = r.winner(), r.depth(), r.z()
    bad = w != o["winner"]
    print("fill", fill, "differing pixels", int(bad.sum()))
    pairs = {}
    for y, x in zip(*np.nonzero(bad)):
        pairs.setdefault((int(w[y, x]), int(o["winner"][y, x])), []).append((x, y))
    ov = o["meshes"][0]["overtex"]
    for (pw, ow), px in sorted(pairs.items(), key=lambda kv: -len(kv[1]))[:6]:
        x, y = px[0]
        print(f" product winner {pw} oracle winner {ow}: {len(px)} px, e.g. ({x},{y}) product depth {d[y,x]} z {z[y,x]} | oracle depth {o['depth'][y,x]} z {o['z'][y,x]} edge_mask {o['edge_mask'][y,x]}")
        for prim in (pw, ow):
            if prim == 0xFFFFFFFF: continue
            idx = sc.meshes[0].indices.reshape(-1, 3)[prim]
            P = ov[idx][:, :4]
            sx = (P[:, 0] + 1) * 0.5 * sc.width; sy = (1 - P[:, 1]) * 0.5 * sc.height
            print("    prim", prim, "screen", np.stack([sx, sy], 1).tolist(), "z", P[:, 2].tolist(), "w", P[:, 3].tolist())
    r.close()
