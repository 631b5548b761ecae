import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_gpgpu_rasterizer_b200 import scenes, _lib
from custom_gpgpu_rasterizer_b200.renderer import Renderer
sc = scenes.sphere_grid_scene(3840, 2160)
r = Renderer(sc.width, sc.height); r.set_option(_lib.OPT_SORT_BINS, 0); r.load_scene(sc); r.render_scene(sc)
counts, offs, ids = r.bin_lists()
nz = np.nonzero((counts > 256) & (counts <= 1024))[0]
rng = np.random.default_rng(0)
for t in rng.choice(nz, 6, replace=False):
    seg = ids[offs[t]:offs[t] + counts[t]].astype(np.int64)
    starts = np.concatenate([[0], np.nonzero(seg[1:] < seg[:-1])[0] + 1, [seg.size]])
    runs = [(seg[a], seg[b - 1], b - a) for a, b in zip(starts[:-1], starts[1:])]
    runs_sorted = sorted(runs)
    bad = sum(1 for x, y in zip(runs_sorted[:-1], runs_sorted[1:]) if x[1] > y[0])
    warps = len(set(seg // 32))
    print(f"tile {t}: n={seg.size} maximal runs={len(runs)} interleaved pairs={bad} distinct warps={warps} first runs={runs[:6]}")
