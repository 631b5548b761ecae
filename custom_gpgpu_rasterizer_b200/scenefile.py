"""Scene description shared by the product harness, the CPU oracle and the reference driver.

A ``Scene`` is plain data: resolution, bin grid, the three render-state switches, the camera
constructor arguments, RGBA8 textures and meshes in the reference's 56-byte ``IVertex`` wire
format.  ``to_blob()`` serialises it to the little-endian ``B2RS`` dump that
``oracle/ref_driver.cu`` (the unmodified reference renderer) and ``oracle/gpu_path_oracle.cpp``
read, so all three sides see byte-identical inputs.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

IVERTEX_FLOATS = 14  # p3 uv2 n3 tan3 c3 (Vertex.h:113-145)
OVERTEX_FLOATS = 18  # p4 uv2 n3 tan3 vd3 c3 (Vertex.h:280-327)


@dataclass
class MeshData:
    vertices: np.ndarray                      # float32 [n, 14]
    indices: np.ndarray                       # uint32 [3 * n_tris]
    position: Sequence[float] = (0.0, 0.0, 0.0)   # Mesh constructor `pos` (Mesh.h:17-20)
    tex_ids: Sequence[int] = (-1, -1, -1, -1)     # diffuse, normal, specular, gloss
    topology: int = 0

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32).reshape(-1, IVERTEX_FLOATS)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32).reshape(-1)

    @property
    def n_tris(self) -> int:
        return self.indices.size // 3


@dataclass
class Scene:
    width: int
    height: int
    meshes: List[MeshData] = field(default_factory=list)
    textures: List[np.ndarray] = field(default_factory=list)   # uint32 [h, w], byte 0 = R
    bins_x: int = 0
    bins_y: int = 0
    bin_w: int = 16
    bin_h: int = 16
    queue_cap: int = 256
    cull: int = 0
    sample: int = 0
    vis: int = 0
    cam_pos: Sequence[float] = (0.0, 1.0, 5.0)   # main.cpp:422
    fov_deg: float = 45.0
    near: float = 0.1
    far: float = 100.0
    pitch_deg: float = 0.0
    yaw_deg: float = 0.0
    name: str = "scene"

    def __post_init__(self):
        if not self.bins_x:
            self.bins_x = (self.width + self.bin_w - 1) // self.bin_w
        if not self.bins_y:
            self.bins_y = (self.height + self.bin_h - 1) // self.bin_h

    @property
    def n_tris(self) -> int:
        return sum(m.n_tris for m in self.meshes)

    @property
    def n_vertices(self) -> int:
        return sum(m.vertices.shape[0] for m in self.meshes)

    def prim_bases(self) -> List[int]:
        out, base = [], 0
        for m in self.meshes:
            out.append(base)
            base += m.n_tris
        return out

    def to_blob(self) -> bytes:
        parts = [struct.pack(
            "<4sI2I5I3I3f5f2II", b"B2RS", 1, self.width, self.height,
            self.bins_x, self.bins_y, self.bin_w, self.bin_h, self.queue_cap,
            self.cull, self.sample, self.vis,
            *[float(v) for v in self.cam_pos],
            self.fov_deg, self.near, self.far, self.pitch_deg, self.yaw_deg,
            len(self.textures), len(self.meshes), 0)]
        for t in self.textures:
            t = np.ascontiguousarray(t, dtype=np.uint32)
            parts.append(struct.pack("<2I", t.shape[1], t.shape[0]))
            parts.append(t.tobytes())
        for m in self.meshes:
            parts.append(struct.pack("<3I3f4i", m.vertices.shape[0], m.indices.size, m.topology,
                                     *[float(v) for v in m.position], *[int(t) for t in m.tex_ids]))
            parts.append(m.vertices.tobytes())
            parts.append(m.indices.tobytes())
        return b"".join(parts)


def read_ref_dump(path: str, scene: Scene) -> dict:
    """Parse the B2RO dump written by oracle/_ref/ref_render (see oracle/ref_driver.cu)."""
    data = np.fromfile(path, dtype=np.uint8)
    off = 0

    def take(dtype, count):
        nonlocal off
        n = np.dtype(dtype).itemsize * count
        out = np.frombuffer(data, dtype=dtype, count=count, offset=off)
        off += n
        return out

    hdr = take(np.uint32, 8)
    assert hdr[0] == 0x4F523242 and hdr[1] == 1, "not a B2RO dump"
    w, h, total, visible, n_meshes, stages = (int(v) for v in hdr[2:8])
    npx = w * h
    out = {"width": w, "height": h, "total_tris": total, "visible_tris": visible}
    out["view_proj"] = take(np.float32, 16).copy()
    out["view"] = take(np.float32, 16).copy()
    out["proj"] = take(np.float32, 16).copy()
    out["cam_pos"] = take(np.float32, 3).copy()
    out["cam_fwd"] = take(np.float32, 3).copy()
    out["framebuffer"] = take(np.uint32, npx).reshape(h, w).copy()
    out["depth"] = take(np.int32, npx).reshape(h, w).copy()
    out["z"] = take(np.float32, npx).reshape(h, w).copy()
    out["frag"] = take(np.float32, npx * 13).reshape(h, w, 13).copy()
    if stages:
        out["meshes"] = []
        for _ in range(n_meshes):
            n_vis, n_pairs = (int(v) for v in take(np.uint32, 2))
            vis = take(np.uint32, n_vis * 3).reshape(-1, 3).copy()
            pairs = take(np.uint32, n_pairs * 4).reshape(-1, 4).copy()
            n_v = int(take(np.uint32, 1)[0])
            ov = take(np.float32, n_v * OVERTEX_FLOATS).reshape(-1, OVERTEX_FLOATS).copy()
            out["meshes"].append({"visible_triples": vis, "bin_pairs": pairs, "overtex": ov})
    return out
