"""Checks of the product at BASELINE.json's FULL sizes (10 M triangles / 64 layers at 3840x2160), where the CPU oracle
cannot render the whole scene in seconds.  They rest on two properties of the path:

* locality - a pixel depends only on the triangles that cover it.  The spheres of config 3 do not overlap on screen, so
  inside the screen rectangle of sphere k the frame of the 10 M-triangle scene must equal, bit for bit, the oracle's
  frame of a scene holding only that sphere's 16 000 triangles (same vertices, same camera); winners differ by the id
  offset k * 16 000;
* depth order - of the 64 stacked quads of config 4 the nearest one wins every pixel, so the frame equals the oracle's
  frame of the nearest layer alone.

The helpers take the product's surfaces as plain arrays, so tests/test_cpu_oracle.py runs the same code on a small grid
with the oracle standing in for the product."""
import dataclasses

import numpy as np

import oracle_lib
from custom_gpgpu_rasterizer_b200.scenefile import MeshData

NONE = 0xFFFFFFFF
INT_MAX = 0x7FFFFFFF
CLEAR = 0x003F3F3F        # default clear colour 0.25 (CUDARenderer.h:135) through the truncating pack
# fragments (inside + z in [0,1]) of BASELINE config 4 at 3840x2160: the oracle's `frags_covered` for the 64-layer scene
# (tests/test_cpu_oracle.py recomputes it when B2R_SLOW_TESTS=1)
CONFIG4_FRAGMENTS = 530667984


def sphere_sub_scene(full, sid: int, nv: int, nt: int):
    """Scene with only sphere `sid` of a sphere_grid_scene (nv vertices / nt triangles per sphere)."""
    m = full.meshes[0]
    v = m.vertices[sid * nv:(sid + 1) * nv]
    i = m.indices[sid * nt * 3:(sid + 1) * nt * 3].astype(np.int64) - sid * nv
    assert i.min() >= 0 and i.max() < nv, "sphere does not own a contiguous vertex range"
    return dataclasses.replace(full, meshes=[MeshData(v, i.astype(np.uint32))], name=f"{full.name}_sphere{sid}")


def nearest_layer_scene(full, layers: int):
    """Scene with only the last (nearest) quad of a fill_quads_scene."""
    m = full.meshes[0]
    v = m.vertices[(layers - 1) * 4:layers * 4]
    i = m.indices[(layers - 1) * 6:layers * 6].astype(np.int64) - (layers - 1) * 4
    return dataclasses.replace(full, meshes=[MeshData(v, i.astype(np.uint32), tex_ids=m.tex_ids)], name=full.name + "_nearest")


def check_sphere_regions(full, fb, depth, winner, sids, nv: int, nt: int, fill_rule: int = 1) -> int:
    """Compare the product's surfaces with the oracle's single-sphere frames inside each sphere's screen rectangle
    (2 pixels of margin).  Returns the number of covered pixels that were compared."""
    compared = 0
    for sid in sids:
        o = oracle_lib.render(sphere_sub_scene(full, sid, nv, nt), fill_rule=fill_rule)
        cov = o["winner"] != NONE
        assert cov.any(), f"sphere {sid} covers no pixel"
        ys, xs = np.nonzero(cov)
        reg = (slice(max(int(ys.min()) - 2, 0), int(ys.max()) + 3), slice(max(int(xs.min()) - 2, 0), int(xs.max()) + 3))
        want_w = o["winner"][reg].astype(np.int64)
        want_w[want_w != NONE] += sid * nt
        assert np.array_equal(winner[reg].astype(np.int64), want_w), f"sphere {sid}: winning primitives differ from the oracle"
        assert np.array_equal(depth[reg], o["depth"][reg]), f"sphere {sid}: integer depth differs from the oracle"
        assert np.array_equal(fb[reg], o["framebuffer"][reg]), f"sphere {sid}: colours differ from the oracle"
        compared += int(cov.sum())
    return compared


def check_whole_frame(o, fb, depth, winner, counts, ids, visible: int):
    """The product's surfaces and lists against the oracle's frame `o` of the same scene: winners, integer depth and
    (untextured scene: no libm on the way) colours bit for bit, the per-tile sets equal, the lists ascending."""
    assert visible == o["visible_tris"], f"visible triangles {visible}, oracle {o['visible_tris']}"
    assert np.array_equal(winner, o["winner"]), f"winning primitive differs at {int((winner != o['winner']).sum())} pixels"
    assert np.array_equal(depth, o["depth"]), f"integer depth differs at {int((depth != o['depth']).sum())} pixels"
    assert np.array_equal(fb, o["framebuffer"]), f"colour differs at {int((fb != o['framebuffer']).sum())} pixels"
    assert np.array_equal(counts, o["bin_counts"]), "per-tile list lengths differ"
    # the oracle's lists are in submission order per tile (ascending already, but do not rely on it): sort segment-wise
    offs = np.concatenate([[0], np.cumsum(o["bin_counts"].astype(np.int64))])
    seg = np.repeat(np.arange(len(o["bin_counts"]), dtype=np.int64), o["bin_counts"].astype(np.int64))
    want = o["bin_ids"].astype(np.int64)[np.lexsort((o["bin_ids"].astype(np.int64), seg))]
    assert offs[-1] == ids.size and np.array_equal(ids.astype(np.int64), want), "tile lists differ from the oracle's sets in ascending order"


def check_frame_invariants(fb, depth, winner):
    empty = winner == NONE
    assert np.array_equal(depth == INT_MAX, empty), "depth is cleared exactly where no primitive won"
    assert (fb[empty] == CLEAR).all(), "untouched pixels must hold the clear colour"


def check_lists(counts, offsets, ids, pairs: int):
    """Exact-sized lists: offsets are the exclusive scan of the counts, every list is in ascending primitive order."""
    counts64 = counts.astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(counts64)])
    assert offs[-1] == ids.size == pairs, f"sum of counts {offs[-1]}, ids {ids.size}, pairs counter {pairs}"
    assert np.array_equal(offsets.astype(np.int64), offs[:-1]), "offsets are not the exclusive scan of the counts"
    if ids.size > 1:
        rising = np.diff(ids.astype(np.int64)) > 0
        starts = offs[1:-1]
        starts = starts[(starts > 0) & (starts < ids.size)]
        rising[starts - 1] = True                     # the step from one list into the next may fall
        assert rising.all(), f"{int((~rising).sum())} list entries are not in ascending primitive order"
