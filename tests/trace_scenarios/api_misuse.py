"""Scenario run by tests/test_cpu_trace.py against the GPU-less tracing build (optionally the AddressSanitizer variant):
the C ABI called the WRONG way - out-of-order calls, bad ids, null pointers, bad sizes, options at the wrong time.  Every
call must come back with a negative status and a message (`int` status everywhere, no exception and no
crash crosses the boundary), and the context must stay usable: a normal frame is rendered in between and at the end.
Prints one line per check; exits non-zero on the first unexpected status."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from custom_gpgpu_rasterizer_b200 import _lib, scenes  # noqa: E402
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera, host_mesh_world  # noqa: E402

lib = _lib.lib
fp = C.POINTER(C.c_float)
checks = 0


def expect(rc, want, what, ctx=None):
    global checks
    checks += 1
    ok = (rc < 0) if want == "error" else (rc == want)
    msg = lib.b2r_last_error(ctx) or b""
    print(f"{'ok  ' if ok else 'FAIL'} {what}: rc {rc}" + (f" ({msg.decode()[:90]})" if rc < 0 else ""))
    if not ok or (rc < 0 and not msg):
        sys.exit(f"unexpected status for: {what}")


sc = scenes.random_triangles_scene(320, 240, 150, seed=5, cull=2, textured=True, n_meshes=2)
cam = host_camera(sc.cam_pos, sc.fov_deg, sc.width, sc.height, sc.near, sc.far)
f = lambda a: a.ctypes.data_as(fp)          # noqa: E731
ident = np.eye(4, dtype=np.float32).reshape(16)
ctx = C.c_void_p()
expect(lib.b2r_create(0, 320, 240, None), "error", "create with a null out pointer")
expect(lib.b2r_create(7, 320, 240, C.byref(ctx)), "error", "create on a device that does not exist")
expect(lib.b2r_create(0, 320, 240, C.byref(ctx)), 0, "create")
h = ctx
buf = np.zeros((240, 320), np.uint32)
out_id = C.c_int(-1)
u64 = C.c_uint64(0)

# ---- before anything was rendered
for name in ("b2r_read_framebuffer", "b2r_read_depth", "b2r_read_z", "b2r_read_winner"):
    expect(getattr(lib, name)(h, buf.ctypes.data), "error", f"{name} before the first frame", h)
expect(lib.b2r_read_bins(h, None, None, None, 0, C.byref(u64)), "error", "read_bins before the first frame", h)
expect(lib.b2r_present_async(h, buf.ctypes.data), "error", "present_async before the first frame", h)
expect(lib.b2r_present_wait(h), 0, "present_wait with nothing pending")
expect(lib.b2r_get_stage_ms(h, (C.c_float * 9)(), 9), "error", "stage times without a timed frame", h)
expect(lib.b2r_render(h, None, None, None, 0, 0, 0, None), "error", "render with null camera", h)
expect(lib.b2r_render(h, f(cam["cam_pos"]), f(cam["cam_fwd"]), f(cam["view_proj"]), 7, 0, 0, None), "error", "render with an unknown cull mode", h)
expect(lib.b2r_render(h, f(cam["cam_pos"]), f(cam["cam_fwd"]), f(cam["view_proj"]), 0, 0, 0, None), 0, "render of an empty scene")
expect(lib.b2r_sync(h), 0, "sync")

# ---- bins, textures, meshes
expect(lib.b2r_setup_bins(h, 0, 0, 16, 16, 256), "error", "setup_bins with no bins", h)
expect(lib.b2r_setup_bins(h, 2, 2, 16, 16, 256), "error", "setup_bins that does not cover the screen", h)
expect(lib.b2r_setup_bins(h, 20, 15, 16, 16, 256), 0, "setup_bins 20 x 15 x 16 x 16")
expect(lib.b2r_texture_create_rgba8(h, None, 4, 4, C.byref(out_id)), "error", "texture with null pixels", h)
px = np.zeros((4, 4), np.uint32)
expect(lib.b2r_texture_create_rgba8(h, px.ctypes.data, 0, 4, C.byref(out_id)), "error", "texture with zero width", h)
expect(lib.b2r_texture_create_rgba8(h, px.ctypes.data, 4, 4, C.byref(out_id)), 0, "texture 4 x 4")
tex = out_id.value
expect(lib.b2r_texture_destroy(h, 99), "error", "destroy of an unknown texture", h)
m = sc.meshes[0]
v = np.ascontiguousarray(m.vertices, np.float32)
i = np.ascontiguousarray(m.indices, np.uint32)
t4 = (C.c_int * 4)(tex, -1, -1, -1)
expect(lib.b2r_mesh_upload(h, None, len(v), i.ctypes.data, i.size, 0, f(ident), t4, C.byref(out_id)), "error", "mesh with null vertices", h)
expect(lib.b2r_mesh_upload(h, v.ctypes.data, len(v), i.ctypes.data, i.size, 0, None, t4, C.byref(out_id)), "error", "mesh with a null world matrix", h)
expect(lib.b2r_mesh_upload(h, v.ctypes.data, len(v), i.ctypes.data, i.size, 1, f(ident), t4, C.byref(out_id)), "error", "TriangleStrip mesh", h)
bad = i.copy(); bad[5] = len(v) + 3
expect(lib.b2r_mesh_upload(h, v.ctypes.data, len(v), bad.ctypes.data, bad.size, 0, f(ident), t4, C.byref(out_id)), "error", "mesh with an index out of range", h)
expect(lib.b2r_mesh_upload(h, v.ctypes.data, 0, i.ctypes.data, 0, 0, f(ident), t4, C.byref(out_id)), 0, "empty mesh")
expect(lib.b2r_mesh_upload(h, v.ctypes.data, len(v), i.ctypes.data, i.size, 0, f(ident), t4, C.byref(out_id)), 0, "mesh")
mesh = out_id.value
expect(lib.b2r_mesh_instance(h, 42, f(ident), None, C.byref(out_id)), "error", "instance of an unknown mesh", h)
expect(lib.b2r_mesh_instance(h, mesh, None, None, C.byref(out_id)), "error", "instance with a null world matrix", h)
expect(lib.b2r_mesh_instance(h, mesh, f(host_mesh_world((1.0, 0.0, 0.0))), None, C.byref(out_id)), 0, "instance")
expect(lib.b2r_mesh_set_world(h, -1, f(ident)), "error", "set_world of mesh -1", h)
expect(lib.b2r_render(h, f(cam["cam_pos"]), f(cam["cam_fwd"]), f(cam["view_proj"]), 2, 1, 0, None), 0, "render")
expect(lib.b2r_read_framebuffer(h, None), "error", "read_framebuffer into null", h)
expect(lib.b2r_read_framebuffer(h, buf.ctypes.data), 0, "read_framebuffer")
expect(lib.b2r_read_z(h, buf.ctypes.data), "error", "read_z without debug surfaces", h)
expect(lib.b2r_read_overtex(h, 99, buf.ctypes.data), "error", "read_overtex of an unknown mesh", h)
expect(lib.b2r_read_visible(h, 99, buf.ctypes.data), "error", "read_visible of an unknown mesh", h)
expect(lib.b2r_read_bins(h, None, None, buf.ctypes.data, 0, C.byref(u64)), 0 if u64.value == 0 else "error", "read_bins with no room for the ids")
# a texture is destroyed while a mesh still names it: the mesh falls back to vertex colours (Mesh texture id -> invalid)
expect(lib.b2r_texture_destroy(h, tex), 0, "destroy of a texture in use")
expect(lib.b2r_render(h, f(cam["cam_pos"]), f(cam["cam_fwd"]), f(cam["view_proj"]), 2, 1, 0, None), 0, "render after its texture was destroyed")
expect(lib.b2r_sync(h), 0, "sync")

# ---- options and multi-GPU calls at the wrong time
expect(lib.b2r_set_option(h, 99, 1), "error", "unknown option", h)
expect(lib.b2r_get_option(h, 99, C.byref(C.c_int())), "error", "unknown option (get)", h)
expect(lib.b2r_get_option(h, _lib.OPT_SORT_BINS, None), "error", "get_option into null", h)
expect(lib.b2r_set_option(h, _lib.OPT_TILE_OWNER_COUNT, 0), "error", "0 tile owners", h)
expect(lib.b2r_set_option(h, _lib.OPT_FRAME_RING, 1), "error", "frame ring of depth 1", h)
expect(lib.b2r_set_option(h, _lib.OPT_FRAME_RING, 9), "error", "frame ring of depth 9", h)
expect(lib.b2r_set_option(h, _lib.OPT_TILE_OWNER_COUNT, 2), 0, "2 tile owners")
expect(lib.b2r_set_option(h, _lib.OPT_TILE_OWNER_RANK, 5), 0, "owner rank 5 of 2 (checked at render)")
expect(lib.b2r_render(h, f(cam["cam_pos"]), f(cam["cam_fwd"]), f(cam["view_proj"]), 2, 1, 0, None), "error", "render with owner rank >= owner count", h)
expect(lib.b2r_set_option(h, _lib.OPT_TILE_OWNER_RANK, 0), 0, "owner rank 0")
expect(lib.b2r_set_option(h, _lib.OPT_TILE_OWNER_COUNT, 1), 0, "1 tile owner")
a64, c64 = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
expect(lib.b2r_shard_bind(h, a64, c64), "error", "shard_bind before shard_prepare", h)
expect(lib.b2r_shard_prepare(h, 0, 1, 0, a64, c64), "error", "shard group of one rank", h)
expect(lib.b2r_shard_prepare(h, 3, 2, 0, a64, c64), "error", "rank 3 of 2", h)
expect(lib.b2r_shard_prepare(h, 0, 2, 0, None, c64), "error", "shard_prepare with a null handle buffer", h)
expect(lib.b2r_shard_release(h), 0, "shard_release with nothing prepared")
expect(lib.b2r_shard_prepare(h, 0, 2, 4096, a64, c64), 0, "shard_prepare")
expect(lib.b2r_shard_prepare(h, 0, 2, 4096, a64, c64), "error", "shard_prepare twice", h)
expect(lib.b2r_set_option(h, _lib.OPT_FRAME_RING, 3), "error", "frame ring changed after shard_prepare", h)
expect(lib.b2r_shard_bind(h, None, c64), "error", "shard_bind with null handles", h)
expect(lib.b2r_shard_release(h), 0, "shard_release")
expect(lib.b2r_packed_size(h, 0, 0, C.byref(u64)), "error", "packed_size for 0 ranks", h)
expect(lib.b2r_unpack_tiles(h, None, 0, 2, None), "error", "unpack_tiles with null buffers", h)
expect(lib.b2r_ipc_export(h, None, None), "error", "ipc_export into a null handle", h)
expect(lib.b2r_ipc_export(h, None, a64), 0, "ipc_export of the context's own framebuffer (null pointer)")
expect(lib.b2r_ipc_open(h, None, C.byref(C.c_void_p())), "error", "ipc_open of a null handle", h)

# ---- the context is still good
expect(lib.b2r_mesh_clear_all(h), 0, "clear_all")
lib.b2r_destroy(h)
with Renderer(sc.width, sc.height) as r:
    r.load_scene(sc)
    r.render_scene(sc)
    r.framebuffer()
print(f"api misuse: {checks} checks passed")
