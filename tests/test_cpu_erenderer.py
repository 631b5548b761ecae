"""CPU: the restatement of the reference's software renderer (oracle/erenderer_oracle.cpp - the timed CPU baseline of
bench.py) is pinned against the reference's OWN Elite::Renderer: CustomCUDARasterizer/ERenderer.cpp compiled unmodified by
oracle/build_ref_erenderer.sh (oracle/_ref/libere_ref.so).

* tests/golden/erenderer_reference.npz (tools/make_golden.py erenderer) holds the reference renderer's outputs for the
  frame cases of tests/golden_cases.py and for BASELINE config 1 (bundled textured vehicle, 640x480): SHA-256 of the
  back buffer and of the depth buffer's float bits, plus three full frames.  The restatement must reproduce them bit
  for bit - this part needs neither /root/reference nor the built library.
* When oracle/_ref/libere_ref.so is present (this container after build(); it travels to the GPU box) the two are also
  compared directly on more and larger scenes."""
import dataclasses
import hashlib
import os

import numpy as np
import pytest

import asset_scenes
import golden_cases
import oracle_lib
from custom_gpgpu_rasterizer_b200 import scenes

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "erenderer_reference.npz"))
CLEAR = 0x40404040            # ClearScreen: memset(64) over the XRGB back buffer (ERenderer.cpp:638-643)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cases():
    cases = dict(golden_cases.frame_cases())
    if asset_scenes.have_assets():
        cases["config1_vehicle_640x480"] = lambda: asset_scenes.vehicle_scene(640, 480, sample=0)
        cases["config1_vehicle_640x480_linear"] = lambda: asset_scenes.vehicle_scene(640, 480, sample=1)
    return cases


@pytest.mark.parametrize("name", sorted(_cases()))
def test_restatement_reproduces_the_reference_renderers_frames(name):
    r = oracle_lib.erenderer_render(_cases()[name]())
    fb, depth = r["framebuffer"], r["depth"].view(np.uint32)
    if name + "_fb" in GOLD.files:        # full surfaces: say where, not only that
        assert np.array_equal(fb, GOLD[name + "_fb"]), f"{int((fb != GOLD[name + '_fb']).sum())} back-buffer pixels differ from the reference renderer"
        assert np.array_equal(depth, GOLD[name + "_depth"]), f"{int((depth != GOLD[name + '_depth']).sum())} depth values differ from the reference renderer"
    assert int((~np.isnan(r["depth"])).sum()) == int(GOLD[name + "_drawn"][0])
    assert (_sha(fb), _sha(depth)) == tuple(GOLD[name + "_sha"]), "frame differs from the reference's Elite::Renderer"
    assert np.array_equal(fb == CLEAR, np.isnan(r["depth"])) or name.startswith("hostile"), "pixels are drawn exactly where depth was written"


def test_golden_file_covers_baseline_config_1():
    assert "config1_vehicle_640x480_sha" in GOLD.files and int(GOLD["config1_vehicle_640x480_drawn"][0]) > 30000


def _turned(sc, pitch, yaw):
    from custom_gpgpu_rasterizer_b200.renderer import host_camera
    fwd = host_camera((0.0, 0.0, 0.0), sc.fov_deg, sc.width, sc.height, sc.near, sc.far, pitch, yaw)["cam_fwd"]
    d = np.float32(np.linalg.norm(sc.cam_pos))
    return dataclasses.replace(sc, pitch_deg=pitch, yaw_deg=yaw, cam_pos=tuple(float(d * f) for f in fwd), name=f"{sc.name}_p{pitch}_y{yaw}")


DIRECT = {
    "soup_640x480": lambda: scenes.random_triangles_scene(640, 480, 400, seed=11, cull=2),
    "soup_backface_3_meshes": lambda: scenes.random_triangles_scene(640, 480, 600, seed=12, cull=0, n_meshes=3),
    "soup_frontface": lambda: scenes.random_triangles_scene(320, 240, 300, seed=13, cull=1, n_meshes=2),
    "soup_textured_point": lambda: scenes.random_triangles_scene(640, 480, 300, seed=14, cull=2, textured=True, n_meshes=3),
    "soup_textured_linear": lambda: scenes.random_triangles_scene(640, 480, 300, seed=15, cull=2, textured=True, n_meshes=3, sample=1),
    "soup_depth_vis": lambda: scenes.random_triangles_scene(320, 240, 200, seed=16, cull=2, vis=1),
    "soup_ragged_resolution": lambda: scenes.random_triangles_scene(333, 217, 250, seed=18, cull=2, textured=True),
    "spheres": lambda: scenes.sphere_grid_scene(960, 544, nx=6, ny=4, stacks=41, slices=60, cam_z=14.0),
    "fill_quads_linear": lambda: scenes.fill_quads_scene(640, 368, layers=8, tex_size=256, sample=1),
    "fill_quads_point": lambda: scenes.fill_quads_scene(640, 368, layers=8, tex_size=256, sample=0),
    "hostile": lambda: scenes.hostile_scene(640, 480, 600, seed=21),
    "hostile_textured": lambda: scenes.hostile_scene(640, 480, 360, seed=22, textured=True),
    "slivers": lambda: scenes.sliver_scene(640, 480, 1500, seed=3),
    "shared_edges": lambda: scenes.shared_edge_scene(),
    # Camera::Pitch / Yaw (Camera.cpp:67-81) with the camera moved back along its new forward vector
    "soup_pitched": lambda: _turned(scenes.random_triangles_scene(320, 240, 300, seed=77, cull=2, textured=True), 10.0, 0.0),
    "soup_yawed": lambda: _turned(scenes.random_triangles_scene(320, 240, 300, seed=78, cull=2), 0.0, 25.0),
    "soup_pitched_and_yawed": lambda: _turned(scenes.random_triangles_scene(320, 240, 300, seed=79, cull=2, textured=True, sample=1), -15.0, 200.0),
}
if asset_scenes.have_assets():
    DIRECT.update({
        "vehicle_1280x720_point": lambda: asset_scenes.vehicle_scene(1280, 720, sample=0),
        "vehicle_1280x720_linear_nocull": lambda: asset_scenes.vehicle_scene(1280, 720, sample=1, cull=2),
        # BASELINE config 5's scene on the CPU path: pitched / yawed camera, 172 meshes with positions.  (With back-face
        # culling the CPU renderer's view-independent cull test removes every triangle of this view - also checked.)
        "vehicles172_view389_frontface": lambda: dataclasses.replace(asset_scenes.vehicle_views_scene(389), cull=1),
        "vehicles172_view389_backface": lambda: asset_scenes.vehicle_views_scene(389),
    })


@pytest.mark.skipif(not oracle_lib.have_erenderer_reference(), reason="oracle/_ref/libere_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(DIRECT))
def test_restatement_equals_the_reference_renderer(name):
    sc = DIRECT[name]()
    port, ref = oracle_lib.erenderer_render(sc), oracle_lib.erenderer_render(sc, reference=True)
    assert np.array_equal(port["framebuffer"], ref["framebuffer"]), f"{int((port['framebuffer'] != ref['framebuffer']).sum())} back-buffer pixels differ"
    assert np.array_equal(port["depth"].view(np.uint32), ref["depth"].view(np.uint32)), "depth buffers differ (float bits)"
    drawn = int((~np.isnan(ref["depth"])).sum())
    # not a vacuous comparison: every scene draws, except the two whose triangles the CPU path's cull test removes entirely
    assert (drawn == 0) if name in ("soup_frontface", "vehicles172_view389_backface") else (drawn > 1000), f"{drawn} pixels drawn"
