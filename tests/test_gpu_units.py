"""-m gpu: unit checks of device-side building blocks."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_hoisted_division_is_bit_exact():
    """w / area with the reciprocal hoisted per triangle (raster_kernel.cuh divByArea) == __fdiv_rn."""
    from custom_gpgpu_rasterizer_b200._lib import lib
    for seed in (1, 2, 3, 4):
        bad = C.c_uint64(123)
        assert lib.b2r_selftest_division(0, 1 << 30, seed, C.byref(bad)) == 0
        assert bad.value == 0, f"{bad.value} of 2^30 quotients differ from div.rn (seed {seed})"


def test_sign_only_cull_matches_the_exact_test():
    """The setup kernel decides back faces from the sign of dot(n, camFwd) whenever that sign is certain and runs the
    reference's arithmetic (normalise with IEEE sqrt + reciprocal, then the dot product: CUDARenderer.cu:760-777) only
    for the rest (geometry_kernels.cuh backFace, proof in the comment there).  2^32 adversarial (normal, camera axis)
    pairs per seed - dot products within a few ulp of zero, magnitudes on both sides of the guard, zero / subnormal /
    NaN / Inf components - in both culling modes: no disagreement, and the short cut must decide most of them."""
    from custom_gpgpu_rasterizer_b200._lib import lib
    n = 1 << 32
    for seed in (1, 2):
        counts = (C.c_uint64 * 2)(123, 0)
        assert lib.b2r_selftest_cull(0, n, seed, counts) == 0
        assert counts[0] == 0, f"{counts[0]} of 2^33 back-face decisions differ from the exact test (seed {seed})"
        assert counts[1] > n // 4, f"the short cut decided only {counts[1]} of 2^32 cases: the test does not exercise it"


@pytest.mark.parametrize("res,grid,tess", [((640, 360), (3, 2), (101, 150)), ((320, 180), (1, 1), (201, 300)), ((1920, 1080), (6, 4), (61, 90))])
def test_list_sort_tiers_give_canonical_lists(res, grid, tess):
    """Every sorter tier (the raster kernel's in-CTA run ordering / bitonic fallback for lists of up to 2048 entries, the long
    radix sort before it, and the stand-alone CTA sort b2r_read_bins uses after frames rendered without the in-frame sort)
    must turn the scatter-order lists into ascending primitive ids without changing their content."""
    import numpy as np
    from custom_gpgpu_rasterizer_b200 import _lib, scenes
    from custom_gpgpu_rasterizer_b200.renderer import Renderer
    sc = scenes.sphere_grid_scene(res[0], res[1], nx=grid[0], ny=grid[1], stacks=tess[0], slices=tess[1], cam_z=6.0 * max(grid))
    r = Renderer(sc.width, sc.height)
    r.load_scene(sc)
    r.set_option(_lib.OPT_SORT_BINS, 0)
    r.render_scene(sc)
    c0, o0, i0 = r.bin_lists()
    fb0 = r.framebuffer().copy()
    r.set_option(_lib.OPT_SORT_BINS, 1)
    r.render_scene(sc)
    c1, o1, i1 = r.bin_lists()
    unsorted_tiles = r.counters()["unsorted_tiles"]
    assert np.array_equal(c0, c1) and np.array_equal(o0, o1)
    assert np.array_equal(r.framebuffer(), fb0), "the image must not depend on the list order"
    tiers = {"cta": 0, "long": 0, "beyond": 0}
    for t in np.flatnonzero(c1):
        seg = i1[o1[t]: o1[t] + c1[t]]
        n = int(c1[t])
        tiers["cta" if n <= 2048 else "long" if n <= 40960 else "beyond"] += 1
        if n <= 40960:
            assert np.all(seg[1:] > seg[:-1]), f"tile {t} ({n} entries) is not in ascending primitive order"
            # B2R_OPT_SORT_BINS = 0 leaves the frame's lists in scatter order, but whoever reads them gets the canonical order
            assert np.array_equal(i0[o0[t]: o0[t] + c0[t]], seg), f"tile {t}: lists read back without the in-frame sort are not canonical"
        assert np.array_equal(np.sort(seg), np.sort(i0[o0[t]: o0[t] + c0[t]])), f"tile {t}: content changed by the sort"
    assert unsorted_tiles == tiers["beyond"]
    assert tiers["cta"] > 0 and (tiers["long"] > 0 or res != (320, 180)), f"scene does not exercise both sorters: {tiers}"
    r.close()
