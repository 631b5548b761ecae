#!/usr/bin/env python
"""Single-GPU diagnostic: per-stage device times of BASELINE config 3 cut down to ny of its 25 rows of spheres (the primitive
slice one rank of a sharded run sets up), same camera, same 4K tile grid - which part of a stage is a fixed cost?
usage: python tools/diag_stage_scaling.py [ny ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from custom_gpgpu_rasterizer_b200 import _lib, scenes
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera

for ny in [int(a) for a in sys.argv[1:]] or [25, 12, 6, 3]:
    sc = scenes.sphere_grid_scene(3840, 2160, nx=25, ny=ny)
    r = Renderer(sc.width, sc.height, device=0)
    r.set_option(_lib.OPT_FRAME_RING, 3)
    r.load_scene(sc)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    prepared = r.prepare_camera(host_camera(sc.cam_pos, sc.fov_deg, sc.width, sc.height, sc.near, sc.far))
    for _ in range(5):
        r.render_prepared(prepared, sc.cull, sc.sample, sc.vis)
    r.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(40):
            r.render_prepared(prepared, sc.cull, sc.sample, sc.vis)
        e1.record(stream)
    torch.cuda.synchronize()
    r.set_option(_lib.OPT_STAGE_TIMERS, 1)
    r.stage_ms_mean(reset=True)
    for _ in range(30):
        r.render_prepared(prepared, sc.cull, sc.sample, sc.vis)
    r.sync()
    stages, n = r.stage_ms_mean(reset=True)
    print(f"ny={ny:2d} prims={sum(m.indices.size // 3 for m in sc.meshes):8d} frame {e0.elapsed_time(e1) / 40 * 1e3:7.1f} us |", {k: round(v * 1e3, 1) for k, v in stages.items() if v}, flush=True)
    r.close()
