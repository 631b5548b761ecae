"""Scenes built from the reference's own bundled assets (BASELINE configs 1, 2 and 5).

The OBJ meshes and PNG textures are copied by oracle/build_ref.sh into the git-ignored directory
oracle/_ref/resources/ (they are the reference's data; they travel to the GPU box with the built
oracle, never into this repository's history).  Everything here is test infrastructure; meshes are
parsed with the product's ObjParser and textures decoded with the product's PNG decoder, both of
which are pinned against the reference parser / byte-exact PNG decoding in tests/test_cpu_host.py.
"""
import os
from functools import lru_cache

import numpy as np

from custom_gpgpu_rasterizer_b200.scenefile import MeshData, Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "oracle", "_ref", "resources")


def have_assets() -> bool:
    return os.path.exists(os.path.join(RES, "vehicle.obj")) and os.path.exists(os.path.join(RES, "vehicle_diffuse.png"))


@lru_cache(maxsize=None)
def _vehicle():
    from custom_gpgpu_rasterizer_b200.renderer import decode_png, parse_obj
    v, i, _ = parse_obj(os.path.join(RES, "vehicle.obj"), invert_y=True)        # main.cpp:169 SetInvertYAxis(true)
    tex = [decode_png(os.path.join(RES, f"vehicle_{k}.png")) for k in ("diffuse", "normal", "specular", "gloss")]
    return v, i, tex


def vehicle_scene(width=1920, height=1080, sample=0, vis=0, cull=0) -> Scene:
    """BASELINE config 2 (config 1 at 640x480): the bundled textured vehicle, camera {0,5,65}, fov 45
    (main.cpp:421), texture slots as wired in main.cpp:172-175."""
    v, i, tex = _vehicle()
    return Scene(width, height, meshes=[MeshData(v, i, (0.0, 0.0, 0.0), (0, 1, 2, 3))], textures=list(tex),
                 cull=cull, sample=sample, vis=vis, cam_pos=(0.0, 5.0, 65.0), fov_deg=45.0, near=0.1, far=100.0,
                 name=f"vehicle_{width}x{height}_s{sample}_v{vis}")


def vehicle_views_scene(view: int, n_instances: int = 172, width=1920, height=1080, n_views=1024) -> Scene:
    """BASELINE config 5: `n_instances` vehicles on a 14x13 lattice in the xz-plane (spacing 45) as separate
    Mesh objects with distinct `pos`, camera `view` of 1024 on a circle: Camera(pos, 45), Pitch(-19.654),
    Yaw(360 i / 1024), pos = 743.3 * forward, near 1 / far 2000."""
    from custom_gpgpu_rasterizer_b200.renderer import host_camera
    v, i, tex = _vehicle()
    meshes = []
    for k in range(n_instances):
        cx, cz = k % 14, k // 14
        meshes.append(MeshData(v, i, ((cx - 6.5) * 45.0, 0.0, (cz - 6.0) * 45.0), (0, 1, 2, 3)))
    pitch, yaw = -19.654, 360.0 * view / n_views
    fwd = host_camera((0.0, 0.0, 0.0), 45.0, width, height, 1.0, 2000.0, pitch, yaw)["cam_fwd"]
    pos = tuple(float(np.float32(743.3) * f) for f in fwd)
    return Scene(width, height, meshes=meshes, textures=list(tex), cull=0, sample=0, vis=0, cam_pos=pos, fov_deg=45.0,
                 near=1.0, far=2000.0, pitch_deg=pitch, yaw_deg=yaw, name=f"vehicles{n_instances}_view{view}")
