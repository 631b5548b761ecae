"""-m gpu: the C++ veneer (reference class surface) driven from a C++ program (examples/headless_main.cpp, the
headless replacement of the reference's main.cpp) produces exactly the frame of the C-ABI path."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from custom_gpgpu_rasterizer_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "custom_gpgpu_rasterizer_b200", "csrc")


@pytest.fixture(scope="module")
def headless(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("veneer") / "headless_main")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(ROOT, "examples", "headless_main.cpp"),
                           "-I" + os.path.join(CSRC, "host"), "-L" + CSRC, "-lb2r", "-Wl,-rpath," + CSRC, "-o", exe])
    return exe


@pytest.mark.parametrize("make", [
    lambda: scenes.random_triangles_scene(640, 480, 300, seed=61, cull=0, textured=True, n_meshes=3, sample=1),
    lambda: scenes.sphere_grid_scene(960, 544, nx=5, ny=3, stacks=41, slices=60, cam_z=12.0),
])
def test_cpp_veneer_matches_c_abi(headless, make):
    from custom_gpgpu_rasterizer_b200.renderer import Renderer
    sc = make()
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "s.b2rs"), os.path.join(td, "o.raw")
        with open(sp, "wb") as f:
            f.write(sc.to_blob())
        out = subprocess.run([headless, sp, op], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        got = np.fromfile(op, dtype=np.uint32).reshape(sc.height, sc.width)
    with Renderer(sc.width, sc.height) as r:
        r.load_scene(sc)
        r.render_scene(sc)
        want = r.framebuffer()
        total, visible = r.stats()
    assert f"triangles {total} visible {visible}" in out.stdout
    assert np.array_equal(got, want)
