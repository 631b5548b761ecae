"""ctypes binding of libb2r.so (the C ABI declared in include/b2r.h).

There is no fallback: if the shared library has not been built (``__graft_entry__.build()`` or
``make -C custom_gpgpu_rasterizer_b200/csrc``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2R_LIBRARY: another build of the same ABI - A/B variants of the product library on a GPU box, or one of the TEST builds
# (tools/trace: recording runtime; tools/emu: CPU execution of the kernels).  The test builds are never the product: the
# default is the in-tree libb2r.so, which needs a CUDA device, and bench.py refuses to measure the emulation build.
LIB_PATH = os.environ.get("B2R_LIBRARY") or os.path.join(_HERE, "csrc", "libb2r.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback for the rendering path)"
    )

lib = C.CDLL(LIB_PATH)

# ---- status codes / enums (include/b2r.h)
B2R_OK = 0
B2R_E_INVALID, B2R_E_CUDA, B2R_E_NO_DEVICE, B2R_E_UNSUPPORTED, B2R_E_STATE = -1, -2, -3, -4, -5
CULL_BACKFACE, CULL_FRONTFACE, CULL_NONE = 0, 1, 2
SAMPLE_POINT, SAMPLE_LINEAR = 0, 1
VIS_PBR, VIS_DEPTH, VIS_NORMAL = 0, 1, 2
TOPOLOGY_TRIANGLE_LIST, TOPOLOGY_TRIANGLE_STRIP = 0, 1
OPT_FILL_RULE_TOP_LEFT, OPT_TILE_OWNER_RANK, OPT_TILE_OWNER_COUNT = 1, 2, 3
OPT_DEBUG_SURFACES, OPT_STAGE_TIMERS, OPT_SORT_BINS, OPT_COUNT_FRAGMENTS, OPT_PACKED_OUTPUT = 4, 5, 6, 7, 8
OPT_OVERLAP_VERTEX = 10
OPT_OVERLAP_FRONT = 11
OPT_FRAME_RING = 12
OPT_SHARD_SLICED_VERTEX = 13
OPT_SHARD_LOCAL_COLOUR = 14
OPT_DIRECT_MAX_AREA = 15
STAGE_NAMES = ("vertex", "setup", "bin_scan", "bin_scatter", "bin_sort", "raster_shade", "total", "exchange", "peer_wait")
NUM_STAGES = 9

# every symbol include/b2r.h declares, with its signature
_vp, _u32, _i32, _u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64
_fp, _u32p, _i32p, _u8p, _u64p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
ABI = {
    "b2r_last_error": (C.c_char_p, [_vp]),
    "b2r_create": (_i32, [_i32, _u32, _u32, C.POINTER(_vp)]),
    "b2r_destroy": (None, [_vp]),
    "b2r_setup_bins": (_i32, [_vp, _u32, _u32, _u32, _u32, _u32]),
    "b2r_texture_create_rgba8": (_i32, [_vp, _vp, _u32, _u32, C.POINTER(C.c_int)]),
    "b2r_texture_destroy": (_i32, [_vp, _i32]),
    "b2r_mesh_upload": (_i32, [_vp, _vp, _u32, _vp, _u32, _i32, _fp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2r_mesh_instance": (_i32, [_vp, _i32, _fp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2r_mesh_set_world": (_i32, [_vp, _i32, _fp]),
    "b2r_mesh_clear_all": (_i32, [_vp]),
    "b2r_set_option": (_i32, [_vp, _i32, _i32]),
    "b2r_get_option": (_i32, [_vp, _i32, C.POINTER(C.c_int)]),
    "b2r_render": (_i32, [_vp, _fp, _fp, _fp, _i32, _i32, _i32, _fp]),
    "b2r_sync": (_i32, [_vp]),
    "b2r_read_framebuffer": (_i32, [_vp, _vp]),
    "b2r_present_async": (_i32, [_vp, _vp]),
    "b2r_present_wait": (_i32, [_vp]),
    "b2r_draw_texture": (_i32, [_vp, _vp, _u32, _u32, _i32, _i32, _i32, _fp]),
    "b2r_host_frame_register": (_i32, [_vp, _vp, _u64]),
    "b2r_present_tiles_async": (_i32, [_vp, _u64]),
    "b2r_present_band_async": (_i32, [_vp, _u64]),
    "b2r_present_rows_async": (_i32, [_vp, _u64, _u32, _u32]),
    "b2r_read_depth": (_i32, [_vp, _vp]),
    "b2r_read_z": (_i32, [_vp, _vp]),
    "b2r_read_winner": (_i32, [_vp, _vp]),
    "b2r_read_overtex": (_i32, [_vp, _i32, _vp]),
    "b2r_read_visible": (_i32, [_vp, _i32, _vp]),
    "b2r_read_bins": (_i32, [_vp, _vp, _vp, _vp, _u64, _u64p]),
    "b2r_get_stats": (_i32, [_vp, _u32p, _u32p]),
    "b2r_get_stage_ms": (_i32, [_vp, _fp, _i32]),
    "b2r_get_stage_ms_mean": (_i32, [_vp, _fp, _i32, _u64p, _i32]),
    "b2r_get_counters": (_i32, [_vp, _u64p, _i32]),
    "b2r_set_stream": (_i32, [_vp, _vp]),
    "b2r_device_surfaces": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "b2r_bind_framebuffer": (_i32, [_vp, _vp]),
    "b2r_packed_size": (_i32, [_vp, _i32, _i32, _u64p]),
    "b2r_unpack_tiles": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "b2r_ipc_export": (_i32, [_vp, _vp, _vp]),
    "b2r_ipc_open": (_i32, [_vp, _vp, C.POINTER(_vp)]),
    "b2r_ipc_close": (_i32, [_vp, _vp]),
    "b2r_shard_prepare": (_i32, [_vp, _i32, _i32, _u64, _vp, _vp]),
    "b2r_shard_bind": (_i32, [_vp, _vp, _vp]),
    "b2r_shard_release": (_i32, [_vp]),
}
for _name, (_res, _args) in ABI.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

# host-side veneer shims (custom_gpgpu_rasterizer_b200/csrc/host/b2r_scene.cpp)
lib.b2r_host_camera.restype = None
lib.b2r_host_camera.argtypes = [_fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, _fp, _fp, _fp]
lib.b2r_host_mesh_world.restype = None
lib.b2r_host_mesh_world.argtypes = [_fp, _i32, _fp]
lib.b2r_host_obj_parse.restype = _vp
lib.b2r_host_obj_parse.argtypes = [C.c_char_p, _i32, _i32]
lib.b2r_host_obj_num_vertices.restype = _u32
lib.b2r_host_obj_num_vertices.argtypes = [_vp]
lib.b2r_host_obj_num_indices.restype = _u32
lib.b2r_host_obj_num_indices.argtypes = [_vp]
lib.b2r_host_obj_vertex_type.restype = _i32
lib.b2r_host_obj_vertex_type.argtypes = [_vp]
lib.b2r_host_obj_vertices.restype = _vp
lib.b2r_host_obj_vertices.argtypes = [_vp]
lib.b2r_host_obj_indices.restype = _vp
lib.b2r_host_obj_indices.argtypes = [_vp]
lib.b2r_host_obj_free.restype = None
lib.b2r_host_obj_free.argtypes = [_vp]
lib.b2r_host_png_decode.restype = _i32
lib.b2r_host_png_decode.argtypes = [C.c_char_p, _u32p, _u32p, _vp, _u64]


class B2RError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b2r error {code}: {message}")
        self.code = code

# test hook: hoisted division of the raster kernel vs __fdiv_rn
lib.b2r_selftest_division.restype = _i32
lib.b2r_selftest_division.argtypes = [_i32, _u64, _u32, _u64p]
# test hook: sign-only back-face test of the setup kernel vs the reference's normalise + dot (counts[0] = disagreements,
# counts[1] = cases the short cut decided)
lib.b2r_selftest_cull.restype = _i32
lib.b2r_selftest_cull.argtypes = [_i32, _u64, _u32, _u64p]
