"""torchrun helper: sort-first composition over N ranks must reproduce the single-GPU frame bit for bit.
usage: python -m torch.distributed.run --nproc-per-node N tools/sort_first_check.py [nccl|p2p|sharded]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from custom_gpgpu_rasterizer_b200 import _lib, scenes
from custom_gpgpu_rasterizer_b200.multi_gpu import ShardedSortFirst, SortFirstGather
from custom_gpgpu_rasterizer_b200.renderer import Renderer, host_camera

mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
test_scenes = [scenes.random_triangles_scene(1000, 600, 500, seed=41, cull=2, textured=True, n_meshes=2),
               scenes.sphere_grid_scene(1920, 1080, nx=8, ny=5, stacks=41, slices=60, cam_z=20.0)]
if mode == "sharded":
    test_scenes.append(scenes.fill_quads_scene(1280, 720, layers=6))     # big rectangles: cooperative binner
    test_scenes.append(scenes.hostile_scene(640, 480, 600, seed=21))     # degenerate / non-finite triangles
for sc in test_scenes:
    r = Renderer(sc.width, sc.height, device=local)
    if rank == 0:
        print(f"sort-first[{mode}] frame ring {r.get_option(_lib.OPT_FRAME_RING)}, sliced vertex stage {r.get_option(_lib.OPT_SHARD_SLICED_VERTEX)}")
    r.load_scene(sc)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    cam = host_camera(sc.cam_pos, sc.fov_deg, sc.width, sc.height, sc.near, sc.far)
    if mode == "sharded":
        g = ShardedSortFirst(r, rank, world)
    else:
        r.set_option(_lib.OPT_TILE_OWNER_COUNT, world)
        r.set_option(_lib.OPT_TILE_OWNER_RANK, rank)
        g = SortFirstGather(r, sc, rank, world, stream, mode=mode)
    for _ in range(5):                      # several frames: exercises the double buffering
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
        g.run()
    r.sync()
    dist.barrier()
    torch.cuda.synchronize()
    lists = None
    got = r.framebuffer() if rank == 0 else None
    if mode == "sharded":
        # the per-tile concatenation of the ranks' lists, in rank order, must be the single-GPU list.  With the direct path
        # the small triangles reach the tile owners as key-buffer blocks, not as list entries: complete lists come from
        # the same frame rendered once more with every triangle binned, on every rank (the composed frame is compared too)
        direct = r.get_option(_lib.OPT_DIRECT_MAX_AREA)
        if direct:
            r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)
            r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
            g.run()
            r.sync()
            dist.barrier()
            torch.cuda.synchronize()
            if rank == 0 and not np.array_equal(r.framebuffer(), got):
                print(f"sort-first[{mode}] world={world} {sc.name}: DIFFERENT frames with / without the direct path")
                ok = False
        counts, offsets, ids = r.bin_lists()
        if direct:
            r.set_option(_lib.OPT_DIRECT_MAX_AREA, direct)
        lists = [None] * world
        dist.all_gather_object(lists, (counts, offsets, ids))
    if rank == 0:
        full = Renderer(sc.width, sc.height, device=local)
        full.set_option(_lib.OPT_FRAME_RING, 0)          # the single-GPU frame of record: serial-capable default scheduler, whatever the environment selects
        full.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)     # ... with every triangle binned
        full.load_scene(sc)
        full.render(cam, sc.cull, sc.sample, sc.vis)
        want = full.framebuffer()
        same = bool(np.array_equal(got, want))
        print(f"sort-first[{mode}] world={world} {sc.name}: {'identical' if same else 'DIFFERENT (%d px)' % int((got != want).sum())}")
        ok &= same
        if lists is not None:
            c1, o1, i1 = full.bin_lists()
            csum = sum(l[0].astype(np.int64) for l in lists)
            same_l = bool(np.array_equal(csum, c1.astype(np.int64)))
            for t in np.flatnonzero(c1)[:: max(1, int(np.count_nonzero(c1)) // 500)]:
                cat = np.concatenate([l[2][l[1][t]: l[1][t] + l[0][t]] for l in lists])
                same_l &= bool(np.array_equal(cat, i1[o1[t]: o1[t] + c1[t]]))
            print(f"sort-first[{mode}] world={world} {sc.name} bin lists: {'identical' if same_l else 'DIFFERENT'}")
            ok &= same_l
        full.close()
    dist.barrier()
    if mode == "sharded":
        # frames in flight with a MOVING camera (every frame differs): each composed frame, presented by rank 0 while
        # the next ones are being rendered, must equal the single-GPU frame of the same camera - catches a stage of
        # one frame reading buffers another frame's stage has already recycled
        cams = [host_camera((sc.cam_pos[0] + 0.11 * k, sc.cam_pos[1] - 0.07 * k, sc.cam_pos[2] + 0.05 * k), sc.fov_deg, sc.width, sc.height, sc.near, sc.far)
                for k in range(10)]
        bufs = [np.zeros((sc.height, sc.width), np.uint32) for _ in cams]
        for cam_k, b in zip(cams, bufs):
            r.render(cam_k, sc.cull, sc.sample, sc.vis, sync=False)
            if rank == 0:
                r.present_async(b)
        if rank == 0:
            r.present_wait()
        else:
            r.sync()
        dist.barrier()
        if rank == 0:
            full = Renderer(sc.width, sc.height, device=local)
            full.set_option(_lib.OPT_FRAME_RING, 0)
            full.load_scene(sc)
            bad = 0
            for cam_k, b in zip(cams, bufs):
                full.render(cam_k, sc.cull, sc.sample, sc.vis)
                bad += 0 if np.array_equal(b, full.framebuffer()) else 1
            full.close()
            print(f"sort-first[{mode}] world={world} {sc.name} moving camera, {len(cams)} frames in flight: {'identical' if bad == 0 else 'DIFFERENT (%d frames)' % bad}")
            ok &= bad == 0
        dist.barrier()
    g.close()
    r.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
