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
