"""torchrun helper: sort-first composition over N ranks must reproduce the single-GPU frame bit for bit.
usage: python -m torch.distributed.run --nproc-per-node N tools/sort_first_check.py [nccl|p2p]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from custom_gpgpu_rasterizer_b200 import _lib, scenes
from custom_gpgpu_rasterizer_b200.multi_gpu import SortFirstGather
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera

mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for sc in (scenes.random_triangles_scene(1000, 600, 500, seed=41, cull=2, textured=True, n_meshes=2),
           scenes.sphere_grid_scene(1920, 1080, nx=8, ny=5, stacks=41, slices=60, cam_z=20.0)):
    r = Renderer(sc.width, sc.height, device=local)
    r.set_option(_lib.OPT_TILE_OWNER_COUNT, world)
    r.set_option(_lib.OPT_TILE_OWNER_RANK, rank)
    r.load_scene(sc)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    cam = host_camera(sc.cam_pos, sc.fov_deg, sc.width, sc.height, sc.near, sc.far)
    g = SortFirstGather(r, sc, rank, world, stream, mode=mode)
    for _ in range(3):                      # several frames: exercises the double buffering
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
        g.run()
    r.sync()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        got = r.framebuffer()
        full = Renderer(sc.width, sc.height, device=local)
        full.load_scene(sc)
        full.render(cam, sc.cull, sc.sample, sc.vis)
        want = full.framebuffer()
        full.close()
        same = bool(np.array_equal(got, want))
        print(f"sort-first[{mode}] world={world} {sc.name}: {'identical' if same else 'DIFFERENT (%d px)' % int((got != want).sum())}")
        ok &= same
    dist.barrier()
    g.close()
    r.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
