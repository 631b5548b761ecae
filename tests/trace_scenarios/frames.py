"""Scenario run by tests/test_cpu_trace.py against the GPU-less tracing build (B2R_LIBRARY = build/trace/libb2r_trace.so):
the public API exercised the way the benchmarks and parity tests use it - frames in flight with a moving camera,
pipelined Present, read-backs, stage timers, unsorted lists canonicalised on read, a world matrix changed and a mesh added
between frames.  Kernels do not run under the recording runtime; what is checked is the ORDER the host code establishes
(tools/trace/stream_order.py).     python frames.py <default|serial|novertex|ringD>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from custom_gpgpu_rasterizer_b200 import _lib, scenes  # noqa: E402
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera, host_mesh_world  # noqa: E402

mode = sys.argv[1]
sc = scenes.random_triangles_scene(320, 240, 150, seed=5, cull=2, textured=True, n_meshes=2)
with Renderer(sc.width, sc.height) as r:
    if mode.startswith("ring"):
        r.set_option(_lib.OPT_FRAME_RING, int(mode[4:]))
    if mode == "serial":
        r.set_option(_lib.OPT_OVERLAP_FRONT, 0)
    if mode == "novertex":
        r.set_option(_lib.OPT_OVERLAP_VERTEX, 0)
    r.set_option(_lib.OPT_DEBUG_SURFACES, 1)
    r.load_scene(sc)
    host = [np.zeros((sc.height, sc.width), np.uint32) for _ in range(2)]

    def cam(i):
        return host_camera((0.1 * i, 0.0, 3.0 + 0.05 * i), sc.fov_deg, sc.width, sc.height, sc.near, sc.far)

    for i in range(11):                                   # more frames than host slots (4) and ring slots
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
        if i % 3 == 2:
            r.present_async(host[(i // 3) & 1])           # pipelined Present: the copy overlaps the following frames
    r.present_wait()
    r.framebuffer(), r.depth(), r.z(), r.winner(), r.bin_lists(), r.stats()
    # a world matrix changes between frames in flight (per-mesh constants are re-uploaded), then a mesh is added
    for i in range(3):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    lib = _lib.lib
    w = host_mesh_world((0.5, 0.0, 0.0))
    r._check(lib.b2r_mesh_set_world(r.handle, 0, w.ctypes.data_as(_lib._fp)))
    for i in range(5):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    m = sc.meshes[0]
    r.add_mesh(m.vertices, m.indices, host_mesh_world((0.0, 1.0, 0.0)), m.tex_ids)
    for i in range(5):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    r.framebuffer()
    # stage timers (serial frames), then frames without the in-frame list sort and lists canonicalised when read
    r.set_option(_lib.OPT_STAGE_TIMERS, 1)
    for i in range(3):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    r.sync()
    r.set_option(_lib.OPT_STAGE_TIMERS, 0)
    r.set_option(_lib.OPT_SORT_BINS, 0)
    for i in range(6):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    r.bin_lists()
    r.set_option(_lib.OPT_SORT_BINS, 1)
    for i in range(3):
        r.render(cam(i), sc.cull, sc.sample, sc.vis, sync=False)
    r.overtex(0, sc.meshes[0].vertices.shape[0]), r.visible(0, sc.meshes[0].n_tris)
print("scenario done:", mode)
