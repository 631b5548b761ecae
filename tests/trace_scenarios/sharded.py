"""Scenario run by tests/test_cpu_trace.py against the GPU-less tracing build: the SHARDED sort-first front end with N
contexts of this one process standing for the N ranks (the recording runtime reports N devices and its IPC handles carry
plain addresses).  Every rank issues the same frames - as the ranks of a real group do, each from its own process; the
order the ranks' calls interleave in here is irrelevant to the happens-before graph.  Kernels do not run; what is
checked is the order the host code and the flag protocol establish (tools/trace/stream_order.py).
    python sharded.py <ranks> <default|serial|ringD|ringDsliced>"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from custom_gpgpu_rasterizer_b200 import _lib, scenes  # noqa: E402
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera  # noqa: E402

world, mode = int(sys.argv[1]), sys.argv[2]
lib = _lib.lib
sc = scenes.random_triangles_scene(320, 240, 1500, seed=7, cull=2, textured=True, n_meshes=2)
ranks = [Renderer(sc.width, sc.height, device=d) for d in range(world)]
for r in ranks:
    if mode.startswith("ring"):
        r.set_option(_lib.OPT_FRAME_RING, int(mode[4]))
        if mode.endswith("sliced"):
            r.set_option(_lib.OPT_SHARD_SLICED_VERTEX, 1)
    if mode == "serial":
        r.set_option(_lib.OPT_OVERLAP_FRONT, 0)
    r.load_scene(sc)
arenas, colours = [], []
for k, r in enumerate(ranks):
    a, c = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
    r._check(lib.b2r_shard_prepare(r.handle, k, world, 1 << 14, a, c))
    arenas.append(bytes(a)), colours.append(bytes(c))
all_arenas = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(arenas))
colour0 = (C.c_ubyte * 64).from_buffer_copy(colours[0])
for r in ranks:
    r._check(lib.b2r_shard_bind(r.handle, all_arenas, colour0))
host_barrier = C.CDLL(os.path.join(os.path.dirname(os.environ["B2R_LIBRARY"]), "libfake_cudart.so")).b2r_trace_host_barrier
host_barrier()                                            # multi_gpu.ShardedSortFirst: dist.barrier() - every mapping exists before anybody renders


def cam(i):
    return host_camera((0.1 * i, 0.0, 3.0 + 0.05 * i), sc.fov_deg, sc.width, sc.height, sc.near, sc.far)


host = [np.zeros((sc.height, sc.width), np.uint32) for _ in range(2)]
for i in range(14):                                       # frames in flight on every rank, moving camera
    for r in ranks:
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    if i % 4 == 3:
        ranks[0].present_async(host[(i // 4) & 1])       # the composed frame is rank 0's
ranks[0].present_wait()
ranks[0].framebuffer()
for r in ranks:
    r.sync()
for r in ranks:
    r.set_option(_lib.OPT_STAGE_TIMERS, 1)
for i in range(3):
    for r in ranks:
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
for r in ranks:
    r.sync()
    r.set_option(_lib.OPT_STAGE_TIMERS, 0)
for i in range(6):
    for r in ranks:
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
ranks[0].framebuffer()
for r in ranks:
    r.sync()
host_barrier()                                            # ShardedSortFirst.close(): nobody unmaps while a peer still renders
for r in ranks:
    r._check(lib.b2r_shard_release(r.handle))
    r.close()
print("scenario done:", world, mode)
