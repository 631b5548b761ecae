"""Deterministic synthetic scenes: the workloads BASELINE.json names and
small scenes for the parity tests.  Everything is generated with numpy from fixed seeds.

Winding convention: the reference only rasterises triangles whose three edge functions are
>= 0 (CUDAROPs.cu:128-145), which on a y-down screen is the counter-clockwise winding and in
NDC (y up) has a positive cross-product z; the same triangles survive BackFace culling with the
default camera forward (0,0,1) (CUDARenderer.cu:760-768).  Generators emit that winding.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

from .scenefile import MeshData, Scene

MASK64 = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def _colour_of(item: int) -> Tuple[float, float, float]:
    h = splitmix64(0x5EED0001 + item)
    return (0.25 + 0.75 * ((h & 0xFFFF) / 65535.0), 0.25 + 0.75 * (((h >> 16) & 0xFFFF) / 65535.0), 0.25 + 0.75 * (((h >> 32) & 0xFFFF) / 65535.0))


def make_vertices(pos, uv=None, n=None, tan=None, col=None) -> np.ndarray:
    pos = np.asarray(pos, np.float32).reshape(-1, 3)
    v = np.zeros((pos.shape[0], 14), np.float32)
    v[:, 0:3] = pos
    if uv is not None:
        v[:, 3:5] = uv
    if n is not None:
        v[:, 5:8] = n
    if tan is not None:
        v[:, 8:11] = tan
    if col is not None:
        v[:, 11:14] = col
    return v


def checker_texture(size: int = 1024, seed: int = 1, cells: int = 16) -> np.ndarray:
    """Procedural RGBA8 'uv grid': coloured cells with a fine gradient (byte 0 = R)."""
    ys, xs = np.mgrid[0:size, 0:size]
    cx, cy = xs * cells // size, ys * cells // size
    rng = np.random.default_rng(seed)
    pal = rng.integers(40, 256, size=(cells * cells, 3), dtype=np.uint32)
    cell = pal[(cx + cy * cells)]
    r = (cell[..., 0] * (160 + (xs % (size // cells)) * 95 // (size // cells)) // 255).astype(np.uint32)
    g = (cell[..., 1] * (160 + (ys % (size // cells)) * 95 // (size // cells)) // 255).astype(np.uint32)
    b = cell[..., 2].astype(np.uint32)
    return (r | (g << 8) | (b << 16) | (255 << 24)).astype(np.uint32)


def noise_texture(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 1 << 32, size=(size, size), dtype=np.uint64).astype(np.uint32)


def normal_texture(size: int, seed: int) -> np.ndarray:
    """Tangent-space normal map: mostly +z with a smooth perturbation."""
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    rng = np.random.default_rng(seed)
    fx, fy = rng.uniform(2, 9, 2)
    nx = 0.35 * np.sin(xs / size * 2 * math.pi * fx)
    ny = 0.35 * np.cos(ys / size * 2 * math.pi * fy)
    nz = np.sqrt(np.maximum(0.0, 1 - nx * nx - ny * ny))
    r = np.clip((nx * 0.5 + 0.5) * 255, 0, 255).astype(np.uint32)
    g = np.clip((ny * 0.5 + 0.5) * 255, 0, 255).astype(np.uint32)
    b = np.clip((nz * 0.5 + 0.5) * 255, 0, 255).astype(np.uint32)
    return (r | (g << 8) | (b << 16) | (255 << 24)).astype(np.uint32)


# --------------------------------------------------------------------------- spheres (config 3)
def uv_sphere(stacks: int, slices: int, radius: float = 1.0):
    """(stacks+1) x (slices+1) parametric grid, seam and poles duplicated; degenerate pole halves dropped."""
    i = np.arange(stacks + 1, dtype=np.float64)
    j = np.arange(slices + 1, dtype=np.float64)
    theta = np.pi * i / stacks
    phi = 2 * np.pi * j / slices
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    x = radius * st * np.cos(phi)[None, :]
    y = radius * ct * np.ones_like(phi)[None, :]
    z = radius * st * np.sin(phi)[None, :]
    pos = np.stack([x, y, z], -1).reshape(-1, 3)
    nrm = pos / radius
    uv = np.stack(np.broadcast_arrays(j[None, :] / slices, i[:, None] / stacks), -1).reshape(-1, 2)
    ii, jj = np.meshgrid(np.arange(stacks), np.arange(slices), indexing="ij")
    v00 = (ii * (slices + 1) + jj).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + slices + 1
    v11 = v10 + 1
    t_a = np.stack([v00, v11, v10], -1)      # lower-left half
    t_b = np.stack([v00, v01, v11], -1)      # upper-right half
    row = ii.reshape(-1)
    t_a = t_a[row != stacks - 1]             # south-pole stack: v10 == v11 in space
    t_b = t_b[row != 0]                      # north-pole stack: v00 == v01 in space
    tris = np.concatenate([t_a, t_b], 0)
    order = np.argsort(np.minimum.reduce(tris, axis=1), kind="stable")   # keep mesh-order locality
    return pos, nrm, uv, tris[order].astype(np.uint32)


def sphere_grid_mesh(nx: int, ny: int, stacks: int, slices: int, pitch=(3.9, 2.2), radius: float = 1.0) -> MeshData:
    pos, nrm, uv, tris = uv_sphere(stacks, slices, radius)
    nv = pos.shape[0]
    verts, idx = [], []
    for sy in range(ny):
        for sx in range(nx):
            sid = sx + sy * nx
            c = np.array([(sx - (nx - 1) / 2) * pitch[0], (sy - (ny - 1) / 2) * pitch[1], 0.0])
            verts.append(make_vertices(pos + c, uv=uv, n=nrm, col=np.broadcast_to(_colour_of(sid), (nv, 3))))
            idx.append(tris + sid * nv)
    return MeshData(np.concatenate(verts, 0), np.concatenate(idx, 0).reshape(-1))


def sphere_grid_scene(width: int = 3840, height: int = 2160, nx: int = 25, ny: int = 25, stacks: int = 81, slices: int = 100,
                      cam_z: float = 68.0, cull: int = 0) -> Scene:
    """BASELINE config 3: 25x25 spheres x 16 000 triangles = 10 000 000 triangles, 5 176 250 vertices."""
    mesh = sphere_grid_mesh(nx, ny, stacks, slices)
    return Scene(width, height, [mesh], [], cull=cull, cam_pos=(0.0, 0.0, cam_z), name=f"spheres_{nx}x{ny}_{stacks}x{slices}_{width}x{height}")


# --------------------------------------------------------------------------- stacked quads (config 4)
def fill_quads_scene(width: int = 3840, height: int = 2160, layers: int = 64, tex_size: int = 1024, sample: int = 1,
                     texture: Optional[np.ndarray] = None, fov_deg: float = 45.0) -> Scene:
    """BASELINE config 4: `layers` view-aligned full-screen textured quads, back to front, NoCulling."""
    tan_half = math.tan(math.radians(fov_deg) / 2)
    aspect = width / height
    p00, p11 = 1.0 / (aspect * tan_half), 1.0 / tan_half
    verts, idx = [], []
    for k in range(layers):
        d = 50.0 - 0.5 * k
        hx, hy = (1.0 + d) / p00, (1.0 + d) / p11     # w_clip = 1 + d
        # screen: TL, BL, BR, TR  (NDC y up -> TL has +y)
        pos = [(-hx, hy, -d), (-hx, -hy, -d), (hx, -hy, -d), (hx, hy, -d)]
        uv = [(0.0, 0.0), (0.0, 0.999), (0.999, 0.999), (0.999, 0.0)]
        verts.append(make_vertices(pos, uv=uv, n=[(0, 0, 1)] * 4, col=[(1, 1, 1)] * 4))
        b = 4 * k
        idx += [b + 0, b + 1, b + 2, b + 0, b + 2, b + 3]   # counter-clockwise on the y-down screen
    tex = texture if texture is not None else checker_texture(tex_size)
    mesh = MeshData(np.concatenate(verts, 0), np.asarray(idx, np.uint32), tex_ids=(0, -1, -1, -1))
    return Scene(width, height, [mesh], [tex], cull=2, sample=sample, cam_pos=(0.0, 0.0, 0.0), fov_deg=fov_deg,
                 name=f"fillquads_{layers}_{width}x{height}")


# --------------------------------------------------------------------------- small test scenes
def random_triangles_scene(width: int, height: int, n_tris: int, seed: int, cull: int = 2, textured: bool = False,
                           sample: int = 0, size: float = 1.5, n_meshes: int = 1, vis: int = 0) -> Scene:
    """Random triangle soup in front of the camera with the rasterised winding, vertex colours (or textures)."""
    rng = np.random.default_rng(seed)
    meshes, textures = [], []
    if textured:
        textures = [checker_texture(256, seed), normal_texture(256, seed + 1), noise_texture(256, seed + 2), noise_texture(256, seed + 3)]
    per = [n_tris // n_meshes + (1 if i < n_tris % n_meshes else 0) for i in range(n_meshes)]
    for mi, nt in enumerate(per):
        centre = np.stack([rng.uniform(-4.2, 4.2, nt), rng.uniform(-2.3, 2.3, nt), rng.uniform(-6.0, 1.5, nt)], -1)
        offs = rng.normal(0.0, size, (nt, 3, 3)) * np.array([1.0, 1.0, 0.4])
        tri = centre[:, None, :] + offs
        # flip to the rasterised winding as seen from the default camera at +z
        a, b = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
        ccw = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) > 0
        flip = ~ccw if mi % 2 == 0 else np.zeros(nt, bool)   # odd meshes keep a mix of both windings
        tri[flip] = tri[flip][:, [0, 2, 1]]
        pos = tri.reshape(-1, 3)
        nrm = rng.normal(0, 1, (nt * 3, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        tan = rng.normal(0, 1, (nt * 3, 3)); tan /= np.linalg.norm(tan, axis=1, keepdims=True)
        verts = make_vertices(pos, uv=rng.uniform(0.02, 0.97, (nt * 3, 2)), n=nrm, tan=tan, col=rng.uniform(0.1, 1.0, (nt * 3, 3)))
        tex = (0, 1, 2, 3) if textured and mi % 3 == 0 else ((0, 1, -1, -1) if textured and mi % 3 == 1 else ((0, -1, -1, -1) if textured else (-1, -1, -1, -1)))
        meshes.append(MeshData(verts, np.arange(nt * 3, dtype=np.uint32), tex_ids=tex))
    return Scene(width, height, meshes, textures, cull=cull, sample=sample, vis=vis, cam_pos=(0.0, 0.0, 5.0),
                 name=f"soup_{n_tris}_{seed}_{width}x{height}")


def hostile_scene(width: int = 640, height: int = 480, n_tris: int = 600, seed: int = 21, textured: bool = False) -> Scene:
    """Inputs the reference handles without clipping or validation: triangles that cross the camera plane
    (w <= 0 on some vertices, NDC flying off to +-inf), vertices exactly on it, huge and tiny coordinates,
    NaN / inf positions, normals and uvs, fully degenerate triangles.  Every stage has to treat them like
    the reference does (saturating / NaN->0 float->int conversions, comparisons that are false for NaN)."""
    rng = np.random.default_rng(seed)
    nt = n_tris
    centre = np.stack([rng.uniform(-5, 5, nt), rng.uniform(-3, 3, nt), rng.uniform(-8.0, 9.0, nt)], -1)   # camera at z = 5
    tri = centre[:, None, :] + rng.normal(0.0, 2.0, (nt, 3, 3))
    k = nt // 12
    tri[0 * k:1 * k, :, 2] = 6.0                                   # w_clip = -z_view + 1 = 0 exactly (Camera.cpp:103-107)
    tri[1 * k:2 * k] *= 1.0e6                                      # huge
    tri[2 * k:3 * k] = centre[2 * k:3 * k, None, :] + rng.normal(0, 1e-6, (k, 3, 3))   # sub-ulp-ish slivers
    tri[3 * k:4 * k, 1] = tri[3 * k:4 * k, 0]                      # two equal vertices
    tri[4 * k:5 * k, 2] = tri[4 * k:5 * k, 0]; tri[4 * k:5 * k, 1] = tri[4 * k:5 * k, 0]   # a point
    tri[5 * k:5 * k + 8, 0, 0] = np.nan
    tri[5 * k + 8:5 * k + 16, 1, 1] = np.inf
    tri[5 * k + 16:5 * k + 24, 2, 2] = -np.inf
    pos = tri.reshape(-1, 3)
    nrm = rng.normal(0, 1, (nt * 3, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    tan = rng.normal(0, 1, (nt * 3, 3))
    nrm[::97] = np.nan; tan[::89] = 0.0
    uv = rng.uniform(-0.3, 1.3, (nt * 3, 2)); uv[::101] = np.nan; uv[::103] = 1.0e9
    col = rng.uniform(0.0, 1.5, (nt * 3, 3)); col[::107] = -1.0; col[::109] = np.nan
    verts = make_vertices(pos, uv=uv, n=nrm, tan=tan, col=col)
    textures = [checker_texture(256, seed), normal_texture(256, seed + 1), noise_texture(256, seed + 2), noise_texture(256, seed + 3)] if textured else []
    mesh = MeshData(verts, np.arange(nt * 3, dtype=np.uint32), tex_ids=(0, 1, 2, 3) if textured else (-1, -1, -1, -1))
    return Scene(width, height, [mesh], textures, cull=2, cam_pos=(0.0, 0.0, 5.0), name=f"hostile_{n_tris}_{seed}_{width}x{height}")


def shared_edge_scene(width: int = 256, height: int = 256) -> Scene:
    """A tessellated quad whose interior edges pass exactly through pixel positions (tests the fill rule)."""
    n = 8
    gx, gy = np.meshgrid(np.linspace(-1.0, 1.0, n + 1), np.linspace(1.0, -1.0, n + 1))
    d = 20.0
    tan_half = math.tan(math.radians(45.0) / 2)
    hx, hy = (1.0 + d) * (width / height) * tan_half * 0.5, (1.0 + d) * tan_half * 0.5
    pos = np.stack([gx * hx, gy * hy, np.full_like(gx, -d)], -1).reshape(-1, 3)
    rng = np.random.default_rng(7)
    verts = make_vertices(pos, col=rng.uniform(0.2, 1.0, (pos.shape[0], 3)))
    idx = []
    for r in range(n):
        for c in range(n):
            v00, v01, v10, v11 = r * (n + 1) + c, r * (n + 1) + c + 1, (r + 1) * (n + 1) + c, (r + 1) * (n + 1) + c + 1
            idx += [v00, v10, v11, v00, v11, v01]
    return Scene(width, height, [MeshData(verts, np.asarray(idx, np.uint32))], [], cull=0, cam_pos=(0.0, 0.0, 0.0), name="shared_edges")


def sliver_scene(width: int = 640, height: int = 480, n_tris: int = 1500, seed: int = 3) -> Scene:
    """Needle-thin, tiny and huge triangles: stresses the conservative candidate-pixel boxes of the raster kernel."""
    rng = np.random.default_rng(seed)
    centre = np.stack([rng.uniform(-4.5, 4.5, n_tris), rng.uniform(-2.5, 2.5, n_tris), rng.uniform(-4.0, 1.0, n_tris)], -1)
    ang = rng.uniform(0, 2 * np.pi, n_tris)
    length = 10.0 ** rng.uniform(-2.5, 1.3, n_tris)          # 0.003 .. 20 world units
    thick = length * 10.0 ** rng.uniform(-5.0, -0.3, n_tris)  # aspect ratios up to 1e5
    d = np.stack([np.cos(ang), np.sin(ang), np.zeros(n_tris)], -1)
    nrm = np.stack([-np.sin(ang), np.cos(ang), np.zeros(n_tris)], -1)
    v0 = centre - d * length[:, None] * 0.5
    v1 = centre + d * length[:, None] * 0.5 + rng.normal(0, 0.02, (n_tris, 3)) * np.array([0, 0, 1.0])
    v2 = centre + nrm * thick[:, None] + d * (rng.uniform(-0.5, 0.5, n_tris) * length)[:, None]
    tri = np.stack([v0, v1, v2], 1)
    a, b = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    ccw = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) > 0
    tri[~ccw] = tri[~ccw][:, [0, 2, 1]]
    verts = make_vertices(tri.reshape(-1, 3), col=rng.uniform(0.2, 1.0, (n_tris * 3, 3)))
    return Scene(width, height, [MeshData(verts, np.arange(n_tris * 3, dtype=np.uint32))], [], cull=2, cam_pos=(0.0, 0.0, 5.0), name=f"slivers_{n_tris}_{seed}")
