"""Python mirror of the reference's renderer interface over the C ABI (libb2r.so).

``Renderer`` plays the role of ``CUDARenderer`` (CUDARenderer.h:31-74): ``load_scene`` =
``LoadScene`` + ``SetupBins``, ``render`` = ``Render``, ``present`` = ``Present``.  Everything that
computes goes through the C ABI; numpy is only used to hold host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import lib, B2RError
from .scenefile import Scene, MeshData, OVERTEX_FLOATS


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def host_camera(pos, fov_deg, width, height, near, far, pitch_deg=0.0, yaw_deg=0.0) -> Dict[str, np.ndarray]:
    """Matrices of the veneer's ``Camera`` for the reference call sequence (main.cpp:421-425)."""
    vp = np.zeros(16, np.float32); view = np.zeros(16, np.float32); proj = np.zeros(16, np.float32)
    cp = np.zeros(3, np.float32); cf = np.zeros(3, np.float32)
    p = np.asarray(pos, np.float32)
    lib.b2r_host_camera(_fptr(p), fov_deg, float(width), float(height), near, far, pitch_deg, yaw_deg,
                        _fptr(vp), _fptr(view), _fptr(proj), _fptr(cp), _fptr(cf))
    return {"view_proj": vp, "view": view, "proj": proj, "cam_pos": cp, "cam_fwd": cf}


def host_mesh_world(pos, vector_ctor: bool = False) -> np.ndarray:
    """World matrix the renderer holds for ``Mesh(..., pos)`` after ``LoadScene`` (CUDAMesh.cpp:16-18)."""
    w = np.zeros(16, np.float32)
    p = np.asarray(pos, np.float32)
    lib.b2r_host_mesh_world(_fptr(p), int(vector_ctor), _fptr(w))
    return w


def parse_obj(path: str, invert_y: bool, pointer_flavour: bool = True):
    """The veneer's ``ObjParser`` -> (vertices float32 [n,14], indices uint32, vertexType)."""
    h = lib.b2r_host_obj_parse(path.encode(), int(invert_y), int(pointer_flavour))
    if not h:
        raise FileNotFoundError(path)
    try:
        nv, ni = lib.b2r_host_obj_num_vertices(h), lib.b2r_host_obj_num_indices(h)
        v = np.ctypeslib.as_array(C.cast(lib.b2r_host_obj_vertices(h), C.POINTER(C.c_float)), shape=(nv * 14,)).copy() if nv else np.zeros(0, np.float32)
        i = np.ctypeslib.as_array(C.cast(lib.b2r_host_obj_indices(h), C.POINTER(C.c_uint32)), shape=(ni,)).copy() if ni else np.zeros(0, np.uint32)
        return v.reshape(-1, 14), i, lib.b2r_host_obj_vertex_type(h)
    finally:
        lib.b2r_host_obj_free(h)


def decode_png(path: str) -> np.ndarray:
    w, h = C.c_uint32(), C.c_uint32()
    if lib.b2r_host_png_decode(path.encode(), C.byref(w), C.byref(h), None, 0) != 0:
        raise ValueError(f"cannot decode {path}")
    out = np.zeros((h.value, w.value), np.uint32)
    lib.b2r_host_png_decode(path.encode(), C.byref(w), C.byref(h), out.ctypes.data, out.size)
    return out


class Renderer:
    """One b2r context (one CUDA device, one stream)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.width, self.height = int(width), int(height)
        self._ctx = C.c_void_p()
        rc = lib.b2r_create(device, self.width, self.height, C.byref(self._ctx))
        if rc != _lib.B2R_OK:
            raise B2RError(rc, (lib.b2r_last_error(None) or b"").decode())
        self.n_meshes = 0
        self.bins = None

    # -- plumbing
    def _check(self, rc: int):
        if rc != _lib.B2R_OK:
            raise B2RError(rc, (lib.b2r_last_error(self._ctx) or b"").decode())

    def close(self):
        if self._ctx:
            lib.b2r_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._ctx

    # -- scene
    def set_option(self, option: int, value: int):
        self._check(lib.b2r_set_option(self._ctx, option, int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int(0)
        self._check(lib.b2r_get_option(self._ctx, option, C.byref(v)))
        return v.value

    def setup_bins(self, bins_x: int, bins_y: int, bin_w: int, bin_h: int, queue_cap: int = 256):
        self._check(lib.b2r_setup_bins(self._ctx, bins_x, bins_y, bin_w, bin_h, queue_cap))
        self.bins = (bins_x, bins_y, bin_w, bin_h)

    def add_texture(self, pixels: np.ndarray) -> int:
        px = np.ascontiguousarray(pixels, np.uint32)
        out = C.c_int(-1)
        self._check(lib.b2r_texture_create_rgba8(self._ctx, px.ctypes.data, px.shape[1], px.shape[0], C.byref(out)))
        return out.value

    def add_mesh(self, vertices: np.ndarray, indices: np.ndarray, world: np.ndarray, tex_ids=(-1, -1, -1, -1), topology: int = 0) -> int:
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 14)
        i = np.ascontiguousarray(indices, np.uint32).reshape(-1)
        w = np.ascontiguousarray(world, np.float32).reshape(16)
        t = (C.c_int * 4)(*[int(x) for x in tex_ids])
        out = C.c_int(-1)
        self._check(lib.b2r_mesh_upload(self._ctx, v.ctypes.data, v.shape[0], i.ctypes.data, i.size, topology, _fptr(w), t, C.byref(out)))
        self.n_meshes += 1
        return out.value

    def add_instance(self, src_mesh: int, world: np.ndarray, tex_ids=None) -> int:
        w = np.ascontiguousarray(world, np.float32).reshape(16)
        t = (C.c_int * 4)(*[int(x) for x in tex_ids]) if tex_ids is not None else None
        out = C.c_int(-1)
        self._check(lib.b2r_mesh_instance(self._ctx, src_mesh, _fptr(w), t, C.byref(out)))
        self.n_meshes += 1
        return out.value

    def set_world(self, mesh: int, world: np.ndarray):
        """Per-frame world matrix of a loaded mesh (SceneGraph::Update, SceneGraph.cpp:30-39): takes effect with the next
        render(); frames already enqueued keep the matrices they were enqueued with."""
        w = np.ascontiguousarray(world, np.float32).reshape(16)
        self._check(lib.b2r_mesh_set_world(self._ctx, mesh, _fptr(w)))

    def clear_meshes(self):
        self._check(lib.b2r_mesh_clear_all(self._ctx))
        self.n_meshes = 0

    def load_scene(self, scene: Scene):
        """LoadScene + SetupBins for a ``Scene`` (CUDARenderer.cu:178-204)."""
        assert scene.width == self.width and scene.height == self.height
        self.setup_bins(scene.bins_x, scene.bins_y, scene.bin_w, scene.bin_h, scene.queue_cap)
        self.clear_meshes()
        for t in getattr(self, "_scene_textures", []):       # the device textures of the previous load_scene
            self._check(lib.b2r_texture_destroy(self._ctx, t))
        tex_map = [self.add_texture(t) for t in scene.textures]
        self._scene_textures = list(tex_map)
        uploaded = {}       # meshes sharing the very same vertex / index memory become instances (b2r_mesh_instance)
        for m in scene.meshes:
            tex = [tex_map[t] if 0 <= t < len(tex_map) else -1 for t in m.tex_ids]
            key = (m.vertices.ctypes.data, m.vertices.shape[0], m.indices.ctypes.data, m.indices.size, m.topology)
            if key in uploaded:
                self.add_instance(uploaded[key], host_mesh_world(m.position), tex)
            else:
                uploaded[key] = self.add_mesh(m.vertices, m.indices, host_mesh_world(m.position), tex, m.topology)

    # -- frame
    def render(self, cam: Dict[str, np.ndarray], cull: int = 0, sample: int = 0, vis: int = 0, clear_rgb=None, sync: bool = True):
        cp = np.ascontiguousarray(cam["cam_pos"], np.float32); cf = np.ascontiguousarray(cam["cam_fwd"], np.float32)
        vp = np.ascontiguousarray(cam["view_proj"], np.float32)
        clr = _fptr(np.ascontiguousarray(clear_rgb, np.float32)) if clear_rgb is not None else None
        self._check(lib.b2r_render(self._ctx, _fptr(cp), _fptr(cf), _fptr(vp), cull, sample, vis, clr))
        if sync:
            self.sync()

    def prepare_camera(self, cam: Dict[str, np.ndarray]):
        """Marshal a camera once for render_prepared(): frame loops that re-use a camera skip the per-call conversions."""
        cp = np.ascontiguousarray(cam["cam_pos"], np.float32); cf = np.ascontiguousarray(cam["cam_fwd"], np.float32)
        vp = np.ascontiguousarray(cam["view_proj"], np.float32)
        return (cp, cf, vp, _fptr(cp), _fptr(cf), _fptr(vp))

    def render_prepared(self, prep, cull: int = 0, sample: int = 0, vis: int = 0):
        """b2r_render with a camera marshalled by prepare_camera(); asynchronous."""
        rc = lib.b2r_render(self._ctx, prep[3], prep[4], prep[5], cull, sample, vis, None)
        if rc != _lib.B2R_OK:
            self._check(rc)

    def render_scene(self, scene: Scene, sync: bool = True):
        cam = host_camera(scene.cam_pos, scene.fov_deg, scene.width, scene.height, scene.near, scene.far, scene.pitch_deg, scene.yaw_deg)
        self.render(cam, scene.cull, scene.sample, scene.vis, sync=sync)
        return cam

    def sync(self):
        self._check(lib.b2r_sync(self._ctx))

    # -- read-backs
    def _surface(self, fn, dtype, out=None):
        if out is None:
            out = np.empty((self.height, self.width), dtype)
        self._check(fn(self._ctx, out.ctypes.data))
        return out

    def framebuffer(self, out=None) -> np.ndarray:
        return self._surface(lib.b2r_read_framebuffer, np.uint32, out)

    def present_async(self, out: np.ndarray):
        """Pipelined Present: start copying the newest frame into `out` (pinned), do not wait."""
        self._check(lib.b2r_present_async(self._ctx, out.ctypes.data))

    def register_host_frame(self, frame: Optional[np.ndarray]):
        """Distributed Present: page-lock and map `frame` (host memory all ranks share, e.g. a /dev/shm mapping; at least
        width*height uint32) for this context's device; None releases it."""
        if frame is None:
            self._check(lib.b2r_host_frame_register(self._ctx, None, 0))
        else:
            self._check(lib.b2r_host_frame_register(self._ctx, frame.ctypes.data, frame.nbytes))

    def present_tiles_async(self, byte_offset: int = 0):
        """Start copying the tiles this rank owns of the newest frame into the registered host frame (+ byte_offset)."""
        self._check(lib.b2r_present_tiles_async(self._ctx, int(byte_offset)))

    def present_band_async(self, byte_offset: int = 0):
        """Sharded front end: start copying this rank's band of rows of the frame composed on rank 0 into the registered host
        frame (+ byte_offset) with this rank's own copy engine / PCIe link."""
        self._check(lib.b2r_present_band_async(self._ctx, int(byte_offset)))

    def present_rows_async(self, byte_offset: int, row_begin: int, row_end: int):
        """present_band_async with the band chosen by the caller: rows [row_begin, row_end) of the frame."""
        self._check(lib.b2r_present_rows_async(self._ctx, int(byte_offset), int(row_begin), int(row_end)))

    def present_wait(self):
        self._check(lib.b2r_present_wait(self._ctx))

    def depth(self) -> np.ndarray:
        return self._surface(lib.b2r_read_depth, np.int32)

    def z(self) -> np.ndarray:
        return self._surface(lib.b2r_read_z, np.float32)

    def winner(self) -> np.ndarray:
        return self._surface(lib.b2r_read_winner, np.uint32)

    def overtex(self, mesh: int, n_vertices: int) -> np.ndarray:
        out = np.empty((n_vertices, OVERTEX_FLOATS), np.float32)
        self._check(lib.b2r_read_overtex(self._ctx, mesh, out.ctypes.data))
        return out

    def visible(self, mesh: int, n_tris: int) -> np.ndarray:
        out = np.empty(n_tris, np.uint8)
        self._check(lib.b2r_read_visible(self._ctx, mesh, out.ctypes.data))
        return out

    def bin_lists(self):
        """(counts, offsets, ids) of the last frame; ids are global primitive ids."""
        total = C.c_uint64(0)
        self._check(lib.b2r_read_bins(self._ctx, None, None, None, 0, C.byref(total)))
        nb = self.bins[0] * self.bins[1]
        counts = np.empty(nb, np.uint32); offsets = np.empty(nb, np.uint32); ids = np.empty(max(total.value, 1), np.uint32)
        self._check(lib.b2r_read_bins(self._ctx, counts.ctypes.data, offsets.ctypes.data, ids.ctypes.data, ids.size, C.byref(total)))
        return counts, offsets, ids[: total.value]

    def stats(self):
        t, v = C.c_uint32(), C.c_uint32()
        self._check(lib.b2r_get_stats(self._ctx, C.byref(t), C.byref(v)))
        return t.value, v.value

    def stage_ms(self) -> Dict[str, float]:
        out = (C.c_float * _lib.NUM_STAGES)()
        self._check(lib.b2r_get_stage_ms(self._ctx, out, _lib.NUM_STAGES))
        return dict(zip(_lib.STAGE_NAMES, [float(x) for x in out]))

    def stage_ms_mean(self, reset: bool = True):
        out = (C.c_float * _lib.NUM_STAGES)()
        frames = C.c_uint64(0)
        self._check(lib.b2r_get_stage_ms_mean(self._ctx, out, _lib.NUM_STAGES, C.byref(frames), int(reset)))
        return dict(zip(_lib.STAGE_NAMES, [float(x) for x in out])), frames.value

    def counters(self) -> Dict[str, int]:
        out = (C.c_uint64 * 8)()
        self._check(lib.b2r_get_counters(self._ctx, out, 8))
        keys = ("pairs", "kernels_launched", "tiles_owned", "num_bins", "unsorted_tiles", "big_triangles", "fragments", "sort_fallback_tiles")
        return dict(zip(keys, [int(x) for x in out]))

    def set_stream(self, stream_ptr: int):
        self._check(lib.b2r_set_stream(self._ctx, C.c_void_p(stream_ptr)))

    def device_surfaces(self):
        fb, d = C.c_void_p(), C.c_void_p()
        self._check(lib.b2r_device_surfaces(self._ctx, C.byref(fb), C.byref(d)))
        return fb.value, d.value

    def bind_framebuffer(self, ptr: Optional[int]):
        self._check(lib.b2r_bind_framebuffer(self._ctx, C.c_void_p(ptr or 0)))
