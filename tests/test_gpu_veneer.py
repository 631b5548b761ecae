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


# ---- the reference's OWN scene-construction source (main.cpp: AddTexture + CreateScenes, extracted verbatim at build time by
# oracle/build_ref_dropin.sh and compiled against the veneer into oracle/_ref/dropin_main) driving the renderer
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin_main")
RES = os.path.join(ROOT, "oracle", "_ref", "resources")


def _asset_scene(graph, width, height):
    """What CreateScenes builds for scene graph `graph` (main.cpp:127-181), rebuilt through the Python harness."""
    from custom_gpgpu_rasterizer_b200.renderer import decode_png, parse_obj
    from custom_gpgpu_rasterizer_b200.scenefile import MeshData, Scene
    if graph == 3:      # bunny: untextured, parser default (no y inversion), camera {0,1,5} (main.cpp:422)
        v, i, _ = parse_obj(os.path.join(RES, "lowpoly_bunny.obj"), invert_y=False)
        return Scene(width, height, meshes=[MeshData(v, i)], cam_pos=(0.0, 1.0, 5.0), name="dropin_bunny")
    if graph == 2:      # TukTuk: diffuse map only, camera {0,5,65} (main.cpp:421)
        v, i, _ = parse_obj(os.path.join(RES, "tuktuk.obj"), invert_y=True)
        return Scene(width, height, meshes=[MeshData(v, i, (0.0, 0.0, 0.0), (0, -1, -1, -1))], textures=[decode_png(os.path.join(RES, "tuktuk.png"))],
                     cam_pos=(0.0, 5.0, 65.0), name="dropin_tuktuk")
    import asset_scenes
    return asset_scenes.vehicle_scene(width, height)          # graph 4: the vehicle with its four maps


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(os.path.join(RES, "vehicle.obj"))), reason="oracle/_ref/dropin_main not built (needs /root/reference at build time)")
@pytest.mark.parametrize("graph,cam,sample", [(3, (0, 1, 5), 0), (2, (0, 5, 65), 0), (4, (0, 5, 65), 0), (4, (0, 5, 65), 1)],
                         ids=["bunny", "tuktuk", "vehicle_point", "vehicle_linear"])
def test_reference_scene_code_on_the_veneer_matches_c_abi(graph, cam, sample, tmp_path):
    from custom_gpgpu_rasterizer_b200.renderer import Renderer
    w, h = 640, 480                                            # Resolution::SD, the reference's default (main.cpp:405)
    os.symlink(RES, tmp_path / "Resources")
    op = str(tmp_path / "o.raw")
    out = subprocess.run([DROPIN, str(tmp_path), "render", str(graph), str(w), str(h), *[str(c) for c in cam], "0", str(sample), op],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = np.fromfile(op, dtype=np.uint32).reshape(h, w)
    sc = _asset_scene(graph, w, h)
    sc.sample = sample
    with Renderer(w, h) as r:
        r.load_scene(sc)
        r.render_scene(sc)
        want = r.framebuffer()
        total, visible = r.stats()
    assert f"triangles {total} visible {visible}" in out.stdout
    assert visible > 0 and np.array_equal(got, want)
