"""CPU: the oracle (oracle/gpu_path_oracle.cpp) against golden vectors PRODUCED BY THE REFERENCE ITSELF
(tests/golden/frame_*.npz = outputs of oracle/_ref/ref_render, the unmodified reference CUDA renderer, run
on a B200 by tools/make_golden.py) plus structural properties of the restatement."""
import os

import numpy as np
import pytest

import golden_cases
import oracle_lib
import parity
from custom_gpgpu_rasterizer_b200 import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_bin_rows(sc, o):
    bases = sc.prim_bases()
    ids, counts = o["bin_ids"], o["bin_counts"]
    bin_of = np.repeat(np.arange(len(counts), dtype=np.uint32), counts)
    mesh_of = np.searchsorted(np.asarray(bases[1:] + [1 << 62]), ids, side="right").astype(np.uint32)
    rows = np.zeros((ids.size, 5), np.uint32)
    rows[:, 0], rows[:, 1] = bin_of, mesh_of
    for mi, m in enumerate(sc.meshes):
        sel = mesh_of == mi
        rows[sel, 2:] = m.indices.reshape(-1, 3)[ids[sel] - bases[mi]]
    return parity.sort_rows(rows)


@pytest.mark.parametrize("name", sorted(golden_cases.frame_cases()))
def test_oracle_reproduces_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"frame_{name}.npz"))
    sc = golden_cases.frame_cases()[name]()
    o = oracle_lib.render(sc, fill_rule=0, model_holes=True)
    # host matrices, vertex stage, visibility, bins: bit-exact
    for k in ("view_proj", "cam_pos", "cam_fwd"):
        assert np.array_equal(parity.bits(g[k]), parity.bits(o[k])), k
    assert int(g["visible_tris"][0]) == o["visible_tris"] and int(g["total_tris"][0]) == o["total_tris"]
    assert np.array_equal(parity.bits(g["overtex0"]), parity.bits(o["meshes"][0]["overtex"])), "VS_Kernel output differs"
    want_vis = parity.sort_rows(parity.triples_of(sc, 0, np.nonzero(o["meshes"][0]["visible"])[0]))
    assert np.array_equal(g["visible0"], want_vis), "TA_Kernel visible set differs"
    assert np.array_equal(g["bins"], _oracle_bin_rows(sc, o)), "TB_Kernel bin sets differ"
    # coverage / depth / colour: identical away from exact-edge fragments and depth ties (races in the reference)
    allowed = o["edge_mask"] != 0
    cov_g, cov_o = g["depth"] != 0x7FFFFFFF, o["depth"] != 0x7FFFFFFF
    assert np.array_equal(cov_g | allowed, cov_o | allowed)
    same = cov_g & cov_o & ~allowed
    assert np.array_equal(g["depth"][same], o["depth"][same]), "integer depth differs"
    assert np.array_equal(parity.bits(g["z"][same]), parity.bits(o["z"][same])), "interpolated z differs"
    rep = parity.assert_colour_parity(g["framebuffer"], o["framebuffer"], allowed, name)
    if not sc.textures:
        assert rep["bad_pixels"] == 0 and (rep["exact_fraction"] == 1.0 or allowed.any())


def test_oracle_is_deterministic():
    sc = scenes.random_triangles_scene(160, 128, 80, seed=3, cull=2, textured=True)
    a, b = oracle_lib.render(sc), oracle_lib.render(sc)
    for k in ("framebuffer", "depth", "winner", "bin_ids"):
        assert np.array_equal(a[k], b[k])


def test_bins_match_numpy_restatement():
    """Bin rectangles recomputed independently in numpy from the oracle's OVertex positions (A3-A5 of SURVEY)."""
    sc = scenes.random_triangles_scene(320, 240, 200, seed=9, cull=2)
    o = oracle_lib.render(sc)
    m = o["meshes"][0]
    ndc = m["overtex"][:, :2].astype(np.float32)
    idx = sc.meshes[0].indices.reshape(-1, 3)
    W, H = np.float32(sc.width), np.float32(sc.height)
    sx = ((ndc[:, 0] + np.float32(1)) * np.float32(0.5)) * W
    sy = ((np.float32(1) - ndc[:, 1]) * np.float32(0.5)) * H
    def s16(v):
        return ((v.astype(np.int64) + 32768) % 65536 - 32768).astype(np.int64)
    def trunc(f):
        return np.clip(np.trunc(f.astype(np.float64)), -2**31, 2**31 - 1).astype(np.int64)
    tx, ty = sx[idx], sy[idx]
    xmin = np.clip(s16(trunc(tx.min(1)) - 1), 0, sc.width); xmax = np.clip(s16(trunc(tx.max(1)) + 1), 0, sc.width)
    ymin = np.clip(s16(trunc(ty.min(1)) - 1), 0, sc.height); ymax = np.clip(s16(trunc(ty.max(1)) + 1), 0, sc.height)
    vis = m["visible"].astype(bool)
    assert np.array_equal(m["bbox"][vis][:, 0], xmin[vis]) and np.array_equal(m["bbox"][vis][:, 1], xmax[vis])
    assert np.array_equal(m["bbox"][vis][:, 2], ymin[vis]) and np.array_equal(m["bbox"][vis][:, 3], ymax[vis])
    sets = oracle_lib.bin_sets(o["bin_counts"], o["bin_ids"])
    for by in range(sc.bins_y):
        for bx in range(sc.bins_x):
            bx0 = np.clip(xmin // 16, 0, sc.bins_x); bx1 = np.clip(xmax // 16, 0, sc.bins_x - 1)
            by0 = np.clip(ymin // 16, 0, sc.bins_y); by1 = np.clip(ymax // 16, 0, sc.bins_y - 1)
            want = np.nonzero(vis & (bx0 <= bx) & (bx <= bx1) & (by0 <= by) & (by <= by1))[0]
            assert np.array_equal(sets[bx + by * sc.bins_x], want.astype(np.uint32))


def test_fill_quads_shared_diagonal_is_drawn_once_under_top_left():
    sc = scenes.fill_quads_scene(160, 96, layers=3, tex_size=128, sample=1)
    o = oracle_lib.render(sc, fill_rule=1)
    o0 = oracle_lib.render(sc, fill_rule=0)
    assert (o["winner"] != 0xFFFFFFFF).all() and (o0["winner"] != 0xFFFFFFFF).all()
    # reference rule: both triangles of a quad produce the pixels exactly on their shared diagonal;
    # the top-left rule hands each of them to one triangle only
    assert o["frags_covered"] <= 3 * 160 * 96 < o0["frags_covered"] + 3 * (160 + 96)
    assert o0["frags_covered"] > o["frags_covered"]
    assert (o["framebuffer"] != o0["framebuffer"]).mean() < 0.02


def test_top_left_rule_is_watertight_on_shared_edges():
    sc = scenes.shared_edge_scene(128, 128)
    o = oracle_lib.render(sc, fill_rule=1)
    inside = o["winner"] != 0xFFFFFFFF
    assert o["frags_covered"] == int(inside.sum()), "every covered pixel is produced by exactly one triangle"
    assert (o["edge_mask"] & 1).any(), "the scene must actually exercise exact-edge pixels"


def test_resolution_holes_are_modelled():
    sc = scenes.random_triangles_scene(100, 70, 60, seed=4, cull=2)
    sc.bins_x, sc.bins_y = 6, 4          # 96x64: inside the screen, like the reference requires
    o = oracle_lib.render(sc, fill_rule=0, model_holes=True)
    assert (o["framebuffer"][64:, :] == 0).all() and (o["framebuffer"][:, 96:] == 0).all()   # never shaded (CUDARenderer.cu:1950)
    assert (o["depth"][64:, :] == 0).all()                                                   # never cleared (:1642)


def test_erenderer_port_agrees_with_gpu_path_on_coverage():
    import ctypes as C
    import subprocess
    lib = os.path.join(oracle_lib.ORACLE_DIR, "liborc_erenderer.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", oracle_lib.ORACLE_DIR, "liborc_erenderer.so"])
    L = C.CDLL(lib)
    L.ere_load.restype = C.c_void_p; L.ere_load.argtypes = [C.c_char_p, C.c_uint64]
    L.ere_render.restype = C.c_double; L.ere_render.argtypes = [C.c_void_p] + [C.c_uint32] * 4
    L.ere_backbuffer.restype = C.c_void_p; L.ere_backbuffer.argtypes = [C.c_void_p]
    L.ere_free.argtypes = [C.c_void_p]
    sc = scenes.sphere_grid_scene(320, 224, nx=4, ny=3, stacks=21, slices=30, cam_z=9.0)
    blob = sc.to_blob()
    S = L.ere_load(blob, len(blob))
    assert L.ere_render(S, 0, 0xFFFFFFFF, 0, 0xFFFFFFFF) > 0
    fb = np.frombuffer((C.c_char * (320 * 224 * 4)).from_address(L.ere_backbuffer(S)), dtype=np.uint32).reshape(224, 320).copy()
    L.ere_free(S)
    o = oracle_lib.render(sc)
    cov_cpu, cov_gpu = fb != 0x40404040, o["winner"] != 0xFFFFFFFF
    assert (cov_cpu != cov_gpu).mean() < 0.002       # same inside test; differences only from fused vs unfused edge functions


def test_fullsize_helpers_on_a_small_grid_with_the_oracle_as_product():
    """tests/fullsize.py (the -m gpu checks at BASELINE's full sizes) exercised on CPU: a 5 x 5 grid of the config-3 spheres at the
    same on-screen size, and an 8-layer config 4, with the oracle's whole-scene frame standing in for the product's."""
    import fullsize
    from custom_gpgpu_rasterizer_b200 import scenes
    full = scenes.sphere_grid_scene(768, 432, nx=5, ny=5, cam_z=68.0 / 5)
    nv, nt = full.n_vertices // 25, full.n_tris // 25
    F = oracle_lib.render(full, fill_rule=1)
    fullsize.check_frame_invariants(F["framebuffer"], F["depth"], F["winner"])
    compared = fullsize.check_sphere_regions(full, F["framebuffer"], F["depth"], F["winner"], [0, 4, 12, 20, 24, 7], nv, nt, fill_rule=1)
    assert compared > 6 * 3000
    counts = F["bin_counts"]
    offsets = np.concatenate([[0], np.cumsum(counts.astype(np.int64))[:-1]]).astype(np.uint32)
    # the oracle's lists hold the same sets; bring them into the product's canonical order first
    ids = np.concatenate([np.sort(s) for s in oracle_lib.bin_sets(counts, F["bin_ids"])] or [np.zeros(0, np.uint32)])
    fullsize.check_lists(counts, offsets, ids, int(counts.sum()))
    fullsize.check_whole_frame(F, F["framebuffer"], F["depth"], F["winner"], counts, ids, F["visible_tris"])
    with pytest.raises(AssertionError):
        bad = ids.copy(); bad[[5, 6]] = bad[[6, 5]]
        fullsize.check_lists(counts, offsets, bad, int(counts.sum()))
    with pytest.raises(AssertionError):
        fullsize.check_whole_frame(F, F["framebuffer"], F["depth"], F["winner"], counts, bad, F["visible_tris"])
    with pytest.raises(AssertionError):
        off = F["depth"].copy(); off[F["winner"] != fullsize.NONE] += 1
        fullsize.check_whole_frame(F, F["framebuffer"], off, F["winner"], counts, ids, F["visible_tris"])
    with pytest.raises(AssertionError):
        wrong = F["framebuffer"].copy(); wrong[F["winner"] != fullsize.NONE] ^= 1
        fullsize.check_sphere_regions(full, wrong, F["depth"], F["winner"], [12], nv, nt, fill_rule=1)
    quads = scenes.fill_quads_scene(320, 180, layers=8, tex_size=64, sample=1)
    Q = oracle_lib.render(quads, fill_rule=1)
    o = oracle_lib.render(fullsize.nearest_layer_scene(quads, 8), fill_rule=1)
    cov = o["winner"] != fullsize.NONE           # where the nearest layer produces a fragment it wins (rounding may cost it a border row)
    assert cov.mean() > 0.98
    assert np.array_equal(Q["framebuffer"][cov], o["framebuffer"][cov]) and np.array_equal(Q["depth"][cov], o["depth"][cov])
    assert np.array_equal(Q["winner"][cov].astype(np.int64), o["winner"][cov].astype(np.int64) + 14)


@pytest.mark.skipif(os.environ.get("B2R_SLOW_TESTS") != "1", reason="~1 min of CPU: the oracle at BASELINE's full sizes (set B2R_SLOW_TESTS=1)")
def test_fullsize_constants_from_the_oracle_at_full_size():
    """The figures tests/test_gpu_zz_fullsize.py relies on, recomputed by the oracle at 3840x2160: the spheres of config 3 are
    disjoint on screen with 2 pixels of margin, all 625 are visible, and config 4 has the fragment count of fullsize.CONFIG4_FRAGMENTS."""
    import fullsize
    from custom_gpgpu_rasterizer_b200 import scenes
    sc = scenes.sphere_grid_scene(3840, 2160)
    F = oracle_lib.render(sc, fill_rule=1)
    assert (F["total_tris"], F["visible_tris"], int(F["bin_counts"].sum())) == (10_000_000, 4_909_458, 7_117_055)
    sids = [0, 24, 600, 624, 12, 300, 324, 612, 312, 137, 401, 538]
    assert fullsize.check_sphere_regions(sc, F["framebuffer"], F["depth"], F["winner"], sids, 8282, 16000, fill_rule=1) > 12 * 4000
    won = F["winner"][F["winner"] != fullsize.NONE]
    assert np.unique(won // 16000).size == 625
    quads = scenes.fill_quads_scene(3840, 2160, layers=64, sample=1)
    Q = oracle_lib.render(quads, fill_rule=1)
    assert Q["frags_covered"] == Q["frags_passed"] == fullsize.CONFIG4_FRAGMENTS
    assert int(Q["bin_counts"].sum()) == 2 * 64 * 240 * 135


class _OracleAsProduct:
    """Stand-in for custom_gpgpu_rasterizer_b200.renderer.Renderer that answers from the CPU oracle: lets the -m gpu test
    FUNCTIONS of tests/test_gpu_zz_fullsize.py run on CPU (their Python, their use of the Renderer API, their constants)."""
    cache = {}

    def __init__(self, width, height, device=0):
        self.opts = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def set_option(self, o, v):
        self.opts[o] = v

    def load_scene(self, scene):
        self.scene = scene

    def render_scene(self, scene, sync=True):
        if scene.name not in self.cache:
            self.cache[scene.name] = oracle_lib.render(scene, fill_rule=1)
        self.o = self.cache[scene.name]

    def framebuffer(self):
        return self.o["framebuffer"]

    def depth(self):
        return self.o["depth"]

    def winner(self):
        return self.o["winner"]

    def stats(self):
        return self.o["total_tris"], self.o["visible_tris"]

    def bin_lists(self):
        counts = self.o["bin_counts"]
        seg = np.repeat(np.arange(len(counts), dtype=np.int64), counts.astype(np.int64))
        ids = self.o["bin_ids"][np.lexsort((self.o["bin_ids"].astype(np.int64), seg))]
        offsets = np.concatenate([[0], np.cumsum(counts.astype(np.int64))[:-1]]).astype(np.uint32)
        return counts, offsets, ids

    def counters(self):
        big = sum(1 for _ in range(0))
        return {"pairs": int(self.o["bin_counts"].sum()), "fragments": int(self.o["frags_covered"]),
                "big_triangles": self.scene.n_tris if self.scene.name.startswith("fillquads") else big}


@pytest.mark.skipif(os.environ.get("B2R_SLOW_TESTS") != "1", reason="~2 min of CPU: the full-size -m gpu test functions with the oracle standing in for the product (set B2R_SLOW_TESTS=1)")
def test_fullsize_gpu_test_functions_run_with_the_oracle_as_product(monkeypatch):
    from custom_gpgpu_rasterizer_b200 import renderer
    import test_gpu_zz_fullsize as t
    monkeypatch.setattr(renderer, "Renderer", _OracleAsProduct)
    t.test_config3_10m_triangles_4k_matches_oracle_per_sphere()
    t.test_config4_64_layers_4k_equals_nearest_layer()
