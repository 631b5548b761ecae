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


def test_veneer_refreshes_moved_meshes_and_draws_textures(headless, tmp_path):
    """SURVEY section 8 f4 through the class surface: a mesh's world matrix changes between two RenderAuto calls (what
    SceneGraph::Update does, SceneGraph.cpp:30-39) and the renderer picks it up (CUDARenderer::Render ->
    b2r_mesh_set_world).  The second frame is the ORACLE's frame of the moved scene: Mesh::m_Position aliases the
    translation of the world matrix (Mesh.h:50), so a shift d of the matrix is a mesh constructed d further along (the
    device's matrix = matrix + position, CUDAMesh.cpp:16-18, doubles both).  Then CUDARenderer::DrawTextureGlobal / DrawTexture (CUDARenderer.h:43-45)
    against a numpy model of the sampler."""
    import copy
    import struct
    import zlib
    import oracle_lib
    sc = scenes.random_triangles_scene(640, 480, 300, seed=67, cull=0, n_meshes=2)
    sc.meshes[0].position = (0.5, 0.25, 0.0)
    d = (0.5, -0.25, 0.125)
    # a 48 x 40 RGBA PNG written by hand (no filter, zlib stored through zlib.compress)
    rng = np.random.default_rng(5)
    tex = rng.integers(0, 256, (40, 48, 4), dtype=np.uint8)
    raw = b"".join(b"\x00" + tex[y].tobytes() for y in range(40))

    def chunk(t, data):
        return struct.pack(">I", len(data)) + t + data + struct.pack(">I", zlib.crc32(t + data))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 48, 40, 8, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    tp = str(tmp_path / "t.png")
    open(tp, "wb").write(png)
    sp, op = str(tmp_path / "s.b2rs"), str(tmp_path / "o.raw")
    open(sp, "wb").write(sc.to_blob())
    out = subprocess.run([headless, sp, op, "--move", *[str(x) for x in d], "--texture", tp, "1", "1", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    first = np.fromfile(op, dtype=np.uint32).reshape(sc.height, sc.width)
    moved = np.fromfile(op + ".moved", dtype=np.uint32).reshape(sc.height, sc.width)
    drawn = np.fromfile(op + ".tex", dtype=np.uint32).reshape(sc.height, sc.width)
    assert np.array_equal(first, oracle_lib.render(sc, fill_rule=1)["framebuffer"])
    sc2 = copy.copy(sc)
    sc2.meshes = [copy.copy(m) for m in sc.meshes]
    sc2.meshes[0].position = tuple(p + x for p, x in zip(sc.meshes[0].position, d))
    want = oracle_lib.render(sc2, fill_rule=1)
    assert not np.array_equal(first, moved)
    assert np.array_equal(moved, want["framebuffer"]), "the frame after the move is not the oracle's frame of the moved scene"
    assert f"moved: triangles {want['total_tris']} visible {want['visible_tris']}" in out.stdout
    # DrawTextureGlobal, stretched, point sampling: CUDATextureSampler::SamplePoint at (x / W, y / H), truncating pack
    # (CUDATextureSampler.cuh:80-104, RGBRaw.h:23-32); fma(u, w, 0.5) = the float64 evaluation rounded once
    W, H = sc.width, sc.height
    u = (np.arange(W, dtype=np.float32) / np.float32(W)).astype(np.float32)
    v = (np.arange(H, dtype=np.float32) / np.float32(H)).astype(np.float32)
    tx = (u.astype(np.float64) * 48 + 0.5).astype(np.float32).astype(np.int64)
    ty = (v.astype(np.float64) * 40 + 0.5).astype(np.float32).astype(np.int64)
    ok = (tx[None, :] < 48) & (ty[:, None] < 40)
    s = tex[np.minimum(ty, 39)[:, None], np.minimum(tx, 47)[None, :]].astype(np.float32)
    c = (s / np.float32(255.0)).astype(np.float32)                      # RGBColor in [0, 1]
    c = np.where(ok[..., None], c[..., :3], np.array([1.0, 0.0, 1.0], np.float32))     # magenta outside
    q = (c * np.float32(255.0)).astype(np.float32).astype(np.uint32) & 255
    want_tex = (q[..., 0] << 16) | (q[..., 1] << 8) | q[..., 2]
    assert np.array_equal(drawn[:H // 16 * 16, :W // 16 * 16], want_tex[:H // 16 * 16, :W // 16 * 16])
