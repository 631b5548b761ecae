"""Sort-first multi-GPU composition (one process per GPU, torch.distributed for the plumbing).

Every rank holds the whole scene, transforms and bins it, but rasterises only the tiles it owns
(owner of tile t = t mod world, t = tx + ty * tiles_x; B2R_OPT_TILE_OWNER_*).  What has to move is
each rank's share of the final colours, W*H*4*(world-1)/world bytes per frame:

* mode "nccl": ranks > 0 render straight into a *packed* tile-major buffer (B2R_OPT_PACKED_OUTPUT),
  one contiguous NCCL gather brings the buffers to rank 0 over NVLink, rank 0 scatters them into
  its row-major framebuffer with one small kernel per source rank (b2r_unpack_tiles).
* mode "p2p" (fused ROP write-back): rank 0 exports its staging buffer through CUDA IPC; the
  raster kernels of the other ranks store their tiles directly into rank 0's memory over NVLink,
  so the transfer overlaps the rasterisation tile by tile and no collective moves pixel data -
  only a 4-byte all-reduce per frame orders "everybody has written" before rank 0 unpacks.

* class ShardedSortFirst (b2r_shard_*): the geometry front end is divided as well - rank r sets up, bins
  and sorts the r-th contiguous slice of the primitives into the lists of ALL tiles; a tile's owner walks
  the ranks' lists in rank order over NVLink and every rank stores its colours straight into rank 0's
  surface.  Small triangles (direct path, B2R_OPT_DIRECT_MAX_AREA) are resolved by the rank that set them up into its own
  key buffer; the touched 16x16 key blocks go to the tiles' owners instead of list entries.  No collective on the data path: two flag barriers per frame inside b2r_render.

The pure index helpers below are shared with the CPU (gloo) tests.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

TILE = 16


def owned_bins(n_bins: int, rank: int, world: int) -> np.ndarray:
    """Bin indices rendered by `rank` (interleaved ownership)."""
    return np.arange(rank, n_bins, world, dtype=np.int64)


def packed_pixels(n_bins: int, rank: int, world: int, sub_tiles: int = 1) -> int:
    return len(owned_bins(n_bins, rank, world)) * sub_tiles * TILE * TILE


def _bin_geometry(width, height, bins_x, bins_y, bin_w, bin_h):
    sub_x, sub_y = (bin_w + TILE - 1) // TILE, (bin_h + TILE - 1) // TILE
    return sub_x, sub_y


def pack_reference(image: np.ndarray, rank: int, world: int, bins_x: int, bins_y: int, bin_w: int = TILE, bin_h: int = TILE) -> np.ndarray:
    """numpy model of B2R_OPT_PACKED_OUTPUT: the packed buffer rank `rank` produces for `image`."""
    h, w = image.shape
    sub_x, sub_y = _bin_geometry(w, h, bins_x, bins_y, bin_w, bin_h)
    bins = owned_bins(bins_x * bins_y, rank, world)
    out = np.zeros(len(bins) * sub_x * sub_y * TILE * TILE, image.dtype)
    for k, b in enumerate(bins):
        bx, by = b % bins_x, b // bins_x
        for s in range(sub_x * sub_y):
            sx, sy = s % sub_x, s // sub_x
            ox, oy = bx * bin_w + sx * TILE, by * bin_h + sy * TILE
            ex, ey = min(ox + TILE, (bx + 1) * bin_w, w), min(oy + TILE, (by + 1) * bin_h, h)
            if ex <= ox or ey <= oy:
                continue
            blk = np.zeros((TILE, TILE), image.dtype)
            blk[: ey - oy, : ex - ox] = image[oy:ey, ox:ex]
            base = (k * sub_x * sub_y + s) * TILE * TILE
            out[base: base + TILE * TILE] = blk.reshape(-1)
    return out


def unpack_reference(packed: np.ndarray, image: np.ndarray, rank: int, world: int, bins_x: int, bins_y: int, bin_w: int = TILE, bin_h: int = TILE) -> None:
    """numpy model of b2r_unpack_tiles: scatter one rank's packed colours into `image` in place."""
    h, w = image.shape
    sub_x, sub_y = _bin_geometry(w, h, bins_x, bins_y, bin_w, bin_h)
    for k, b in enumerate(owned_bins(bins_x * bins_y, rank, world)):
        bx, by = b % bins_x, b // bins_x
        for s in range(sub_x * sub_y):
            sx, sy = s % sub_x, s // sub_x
            ox, oy = bx * bin_w + sx * TILE, by * bin_h + sy * TILE
            ex, ey = min(ox + TILE, (bx + 1) * bin_w, w), min(oy + TILE, (by + 1) * bin_h, h)
            if ex <= ox or ey <= oy:
                continue
            base = (k * sub_x * sub_y + s) * TILE * TILE
            blk = packed[base: base + TILE * TILE].reshape(TILE, TILE)
            image[oy:ey, ox:ex] = blk[: ey - oy, : ex - ox]


def gather_packed(local, rank: int, world: int, dst: int = 0):
    """dist.gather of equally sized per-rank buffers to `dst` (works with nccl and gloo tensors)."""
    import torch
    import torch.distributed as dist
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst)
    return bufs


def prim_slice(n_prims: int, rank: int, world: int, block: int = 256):
    """[lo, hi) of the primitives rank `rank` sets up and bins in the sharded front end (whole 256-blocks)."""
    blocks = (n_prims + block - 1) // block
    lo = min(n_prims, blocks * rank // world * block)
    hi = min(n_prims, blocks * (rank + 1) // world * block)
    return lo, hi


def check_shard_modes(modes) -> None:
    """Every rank of a sharded group must run the same frame-ring depth, vertex-stage mode and direct-path switch (the
    layout of the shared arena depends on them)."""
    if any(m != modes[0] for m in modes):
        raise RuntimeError(f"ranks disagree on (B2R_OPT_FRAME_RING, B2R_OPT_SHARD_SLICED_VERTEX, B2R_OPT_DIRECT_MAX_AREA > 0): {list(modes)}")


class ShardedSortFirst:
    """Sharded front end: exchange the CUDA-IPC handles once; afterwards every Renderer.render() on every
    rank is one frame of the distributed pipeline and the composed image lives on rank 0."""

    launches_per_frame = 0      # the barrier kernels are counted by b2r_render itself

    def __init__(self, renderer, rank: int, world: int, list_capacity: int = 0):
        import torch.distributed as dist
        from . import _lib
        self.dist, self._lib, self.r, self.rank, self.world = dist, _lib, renderer, rank, world
        arena, colour = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
        renderer._check(_lib.lib.b2r_shard_prepare(renderer.handle, rank, world, list_capacity, arena, colour))
        handles: List[Optional[tuple]] = [None] * world
        # the layout of the shared arena depends on the frame-ring depth, the raster kernel on the sliced vertex stage
        mode = (renderer.get_option(_lib.OPT_FRAME_RING), renderer.get_option(_lib.OPT_SHARD_SLICED_VERTEX), int(renderer.get_option(_lib.OPT_DIRECT_MAX_AREA) > 0))
        dist.all_gather_object(handles, (bytes(arena), bytes(colour), mode))
        check_shard_modes([h[2] for h in handles])
        all_arenas = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(h[0] for h in handles))
        colour0 = (C.c_ubyte * 64).from_buffer_copy(handles[0][1])
        renderer._check(_lib.lib.b2r_shard_bind(renderer.handle, all_arenas, colour0))
        dist.barrier()            # every mapping exists before anybody renders
        self.bound = True

    def run(self):
        pass

    def close(self):
        if self.bound:
            self.r.sync()
            self.dist.barrier()   # nobody still reads this rank's arena
            self.r._check(self._lib.lib.b2r_shard_release(self.r.handle))
            self.dist.barrier()
            self.bound = False


class SortFirstGather:
    """Per-frame composition step; call run() right after Renderer.render(sync=False) on every rank."""

    def __init__(self, renderer, scene, rank: int, world: int, stream, mode: str = "nccl"):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.r, self.rank, self.world, self.stream, self.mode = renderer, rank, world, stream, mode
        n_bins = scene.bins_x * scene.bins_y
        sub = ((scene.bin_w + TILE - 1) // TILE) * ((scene.bin_h + TILE - 1) // TILE)
        self.sizes = [packed_pixels(n_bins, k, world, sub) for k in range(world)]
        self.max_px = max(self.sizes)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.launches_per_frame = 0
        self.peer_ptr = None
        if mode == "p2p":
            # rank 0 owns one staging area with a slot per rank and shares it through CUDA IPC
            # two generations: a rank may start writing frame i+1 while rank 0 still unpacks frame i
            self.staging = torch.zeros((2, world, self.max_px), dtype=torch.int32, device=dev) if rank == 0 else None
            self.frame = 0
            handle = (C.c_ubyte * 64)()
            if rank == 0:
                renderer._check(_lib.lib.b2r_ipc_export(renderer.handle, C.c_void_p(self.staging.data_ptr()), handle))
            obj = [bytes(handle)]
            dist.broadcast_object_list(obj, src=0)
            self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
            if rank != 0:
                ptr = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(obj[0])
                renderer._check(_lib.lib.b2r_ipc_open(renderer.handle, hb, C.byref(ptr)))
                self.peer_ptr = ptr.value
                renderer.set_option(_lib.OPT_PACKED_OUTPUT, 1)
                renderer.bind_framebuffer(self._slot(0))
            self.launches_per_frame = 1 + (world - 1 if rank == 0 else 0)
        else:
            self.local = torch.zeros(self.max_px, dtype=torch.int32, device=dev)
            self.bufs = [torch.zeros(self.max_px, dtype=torch.int32, device=dev) for _ in range(world)] if rank == 0 else None
            if rank != 0:
                renderer.set_option(_lib.OPT_PACKED_OUTPUT, 1)
                renderer.bind_framebuffer(self.local.data_ptr())
            self.launches_per_frame = 1 + (world - 1 if rank == 0 else 0)

    def _slot(self, gen: int) -> int:
        return self.peer_ptr + ((gen * self.world + self.rank) * self.max_px) * 4

    def run(self):
        torch, dist = self.torch, self.dist
        with torch.cuda.stream(self.stream):
            if self.mode == "p2p":
                gen = self.frame & 1
                dist.all_reduce(self.flag)          # orders "all ranks' stores have landed" before the unpack
                if self.rank == 0:
                    for k in range(1, self.world):
                        self.r._check(self._lib.lib.b2r_unpack_tiles(self.r.handle, C.c_void_p(self.staging[gen, k].data_ptr()), k, self.world, None))
                else:
                    self.r.bind_framebuffer(self._slot(gen ^ 1))   # next frame writes the other generation
                self.frame += 1
            else:
                dist.gather(self.local, self.bufs, dst=0)
                if self.rank == 0:
                    for k in range(1, self.world):
                        self.r._check(self._lib.lib.b2r_unpack_tiles(self.r.handle, C.c_void_p(self.bufs[k].data_ptr()), k, self.world, None))

    def close(self):
        if self.peer_ptr:
            self.r.bind_framebuffer(None)
            self.r._check(self._lib.lib.b2r_ipc_close(self.r.handle, C.c_void_p(self.peer_ptr)))
            self.peer_ptr = None
