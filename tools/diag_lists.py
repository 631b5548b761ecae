import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_gpgpu_rasterizer_b200 import scenes
from custom_gpgpu_rasterizer_b200.renderer import Renderer
sc = scenes.sphere_grid_scene(3840, 2160)
r = Renderer(sc.width, sc.height); r.load_scene(sc); r.render_scene(sc)
counts, offs, ids = r.bin_lists()
print("tiles", counts.size, "nonempty", (counts > 0).sum(), "pairs", counts.sum(), "max", counts.max())
for lo, hi in [(1, 32), (33, 256), (257, 1024), (1025, 4096), (4097, 40960), (40961, 1 << 30)]:
    m = (counts >= lo) & (counts <= hi)
    print(f"{lo:>6}..{hi:<10} tiles {m.sum():>6} entries {counts[m].sum():>9}")
print(r.counters())
top = np.sort(counts)[::-1][:24]
print("longest lists", top.tolist())
print("tiles > 1024:", int((counts > 1024).sum()), " > 2048:", int((counts > 2048).sum()))
