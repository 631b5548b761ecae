"""Scenario run by tests/test_cpu_trace.py against the GPU-less tracing build: degenerate work sizes - an empty scene, and
scenes with fewer 256-primitive blocks than ranks, so that some ranks own an EMPTY primitive slice (and, with the sliced
vertex stage, no vertex chunk).  Checked: no launch with a zero-sized grid (an error on a real device that the recording
runtime would not raise) and no unordered conflicting accesses.     python tiny.py <ranks> <default|ringD[sliced]> <triangles>"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from custom_gpgpu_rasterizer_b200 import scenes, _lib
from custom_gpgpu_rasterizer_b200.renderer import Renderer
from custom_gpgpu_rasterizer_b200.scenefile import Scene
world, mode, ntris = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
lib = _lib.lib
sc = scenes.random_triangles_scene(320, 240, ntris, seed=7, cull=2) if ntris else Scene(320, 240)
ranks = [Renderer(sc.width, sc.height, device=d) for d in range(world)]
for r in ranks:
    if mode.startswith('ring'):
        r.set_option(_lib.OPT_FRAME_RING, int(mode[4]))
        if mode.endswith('sliced'): r.set_option(_lib.OPT_SHARD_SLICED_VERTEX, 1)
    r.load_scene(sc)
if world > 1:
    arenas, colours = [], []
    for k, r in enumerate(ranks):
        a, c = (C.c_ubyte*64)(), (C.c_ubyte*64)()
        r._check(lib.b2r_shard_prepare(r.handle, k, world, 1<<12, a, c)); arenas.append(bytes(a)); colours.append(bytes(c))
    aa = (C.c_ubyte*(64*world)).from_buffer_copy(b''.join(arenas)); c0 = (C.c_ubyte*64).from_buffer_copy(colours[0])
    for r in ranks: r._check(lib.b2r_shard_bind(r.handle, aa, c0))
for i in range(6):
    for r in ranks: r.render_scene(sc, sync=False)
ranks[0].framebuffer()
for r in ranks: r.sync()
print('done', world, mode, ntris)
