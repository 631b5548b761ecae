"""ctypes binding of the CPU oracle (oracle/liborc_gpu_path.so).  TEST INFRASTRUCTURE: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liborc_gpu_path.so")


def _load():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liborc_gpu_path.so"])
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.orc_render.restype = vp
    lib.orc_render.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int]
    lib.orc_free.argtypes = [vp]
    lib.orc_error.restype = C.c_char_p
    lib.orc_error.argtypes = [vp]
    for name in ("orc_width", "orc_height", "orc_num_meshes", "orc_num_bins", "orc_total_tris", "orc_total_visible"):
        getattr(lib, name).restype = C.c_uint32
        getattr(lib, name).argtypes = [vp]
    for name in ("orc_frags_covered", "orc_frags_passed", "orc_bin_total"):
        getattr(lib, name).restype = C.c_uint64
        getattr(lib, name).argtypes = [vp]
    fp = C.POINTER(C.c_float)
    lib.orc_matrices.argtypes = [vp, fp, fp, fp, fp, fp]
    lib.orc_mesh_world.argtypes = [vp, C.c_uint32, fp]
    for name in ("orc_mesh_num_vertices", "orc_mesh_num_tris", "orc_mesh_prim_base"):
        getattr(lib, name).restype = C.c_uint32
        getattr(lib, name).argtypes = [vp, C.c_uint32]
    for name in ("orc_mesh_overtex", "orc_mesh_visible", "orc_mesh_bbox"):
        getattr(lib, name).restype = vp
        getattr(lib, name).argtypes = [vp, C.c_uint32]
    for name in ("orc_bin_counts", "orc_bin_ids", "orc_framebuffer", "orc_depth", "orc_zbuf", "orc_winner", "orc_edge_mask"):
        getattr(lib, name).restype = vp
        getattr(lib, name).argtypes = [vp]
    lib.orc_frag_attrs.argtypes = [vp, fp]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _arr(ptr, dtype, shape):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape).copy()


def render(scene, fill_rule: int = 0, model_holes: bool = False) -> dict:
    """Run the CPU oracle on a Scene; returns every parity surface as numpy arrays."""
    L = lib()
    blob = scene.to_blob()
    f = L.orc_render(blob, len(blob), int(fill_rule), int(model_holes))
    try:
        err = L.orc_error(f)
        if err:
            raise RuntimeError(err.decode())
        w, h = L.orc_width(f), L.orc_height(f)
        out = {"width": w, "height": h, "total_tris": L.orc_total_tris(f), "visible_tris": L.orc_total_visible(f),
               "frags_covered": L.orc_frags_covered(f), "frags_passed": L.orc_frags_passed(f)}
        vp = np.zeros(16, np.float32); view = np.zeros(16, np.float32); proj = np.zeros(16, np.float32)
        cp = np.zeros(3, np.float32); cf = np.zeros(3, np.float32)
        fp = C.POINTER(C.c_float)
        L.orc_matrices(f, *[a.ctypes.data_as(fp) for a in (vp, view, proj, cp, cf)])
        out.update(view_proj=vp, view=view, proj=proj, cam_pos=cp, cam_fwd=cf)
        meshes = []
        for m in range(L.orc_num_meshes(f)):
            nv, nt = L.orc_mesh_num_vertices(f, m), L.orc_mesh_num_tris(f, m)
            world = np.zeros(16, np.float32)
            L.orc_mesh_world(f, m, world.ctypes.data_as(fp))
            meshes.append({
                "overtex": _arr(L.orc_mesh_overtex(f, m), np.float32, (nv, 18)),
                "visible": _arr(L.orc_mesh_visible(f, m), np.uint8, (nt,)),
                "bbox": _arr(L.orc_mesh_bbox(f, m), np.int16, (nt, 4)),
                "prim_base": L.orc_mesh_prim_base(f, m), "world": world})
        out["meshes"] = meshes
        nb = L.orc_num_bins(f)
        out["bin_counts"] = _arr(L.orc_bin_counts(f), np.uint32, (nb,))
        out["bin_ids"] = _arr(L.orc_bin_ids(f), np.uint32, (int(L.orc_bin_total(f)),))
        out["framebuffer"] = _arr(L.orc_framebuffer(f), np.uint32, (h, w))
        out["depth"] = _arr(L.orc_depth(f), np.int32, (h, w))
        out["z"] = _arr(L.orc_zbuf(f), np.float32, (h, w))
        out["winner"] = _arr(L.orc_winner(f), np.uint32, (h, w))
        out["edge_mask"] = _arr(L.orc_edge_mask(f), np.uint8, (h, w))
        frag = np.zeros((h, w, 13), np.float32)
        L.orc_frag_attrs(f, frag.ctypes.data_as(fp))
        out["frag"] = frag
        return out
    finally:
        L.orc_free(f)


def bin_sets(counts, ids):
    """Per-bin sorted id arrays from (counts, concatenated ids)."""
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return [np.sort(ids[offs[b]:offs[b + 1]]) for b in range(len(counts))]


# ---------------------------------------------------------------------------------------- the CPU software path
ERE_PORT = os.path.join(ORACLE_DIR, "liborc_erenderer.so")          # restatement (oracle/erenderer_oracle.cpp)
ERE_REF = os.path.join(ORACLE_DIR, "_ref", "libere_ref.so")          # the reference's own Elite::Renderer (oracle/build_ref_erenderer.sh)
_ere = {}


def have_erenderer_reference() -> bool:
    return os.path.exists(ERE_REF)


def erenderer_lib(reference: bool = False):
    """ctypes handle of the CPU-path renderer: the restatement, or the reference's own ERenderer.cpp build."""
    path = ERE_REF if reference else ERE_PORT
    if path not in _ere:
        if not reference and not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liborc_erenderer.so"])
        L = C.CDLL(path)
        L.ere_load.restype = C.c_void_p
        L.ere_load.argtypes = [C.c_char_p, C.c_uint64]
        L.ere_error.restype = C.c_char_p
        L.ere_error.argtypes = [C.c_void_p]
        L.ere_render.restype = C.c_double
        L.ere_render.argtypes = [C.c_void_p] + [C.c_uint32] * 4
        L.ere_free.argtypes = [C.c_void_p]
        L.ere_backbuffer.restype = C.c_void_p
        L.ere_backbuffer.argtypes = [C.c_void_p]
        L.ere_depth.restype = C.c_void_p
        L.ere_depth.argtypes = [C.c_void_p]
        _ere[path] = L
    return _ere[path]


def erenderer_render(scene, reference: bool = False) -> dict:
    """One frame of Elite::Renderer::Render (ERenderer.cpp:42-164): 0x00RRGGBB back buffer, float depth buffer (NaN where
    nothing was drawn: the reference clears it with memset(0xFF)), seconds spent inside Render."""
    L = erenderer_lib(reference)
    blob = scene.to_blob()
    S = L.ere_load(blob, len(blob))
    try:
        err = L.ere_error(S)
        if err:
            raise RuntimeError(err.decode())
        seconds = L.ere_render(S, 0, 0xFFFFFFFF, 0, 0xFFFFFFFF)
        if seconds < 0:
            raise RuntimeError("ere_render failed")
        h, w = scene.height, scene.width
        return {"framebuffer": _arr(L.ere_backbuffer(S), np.uint32, (h, w)), "depth": _arr(L.ere_depth(S), np.float32, (h, w)), "seconds": seconds}
    finally:
        L.ere_free(S)
