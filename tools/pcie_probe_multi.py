#!/usr/bin/env python
"""How much device->host bandwidth do N ranks get TOGETHER (torchrun, one rank per GPU)?  Every rank copies its band of a 4K frame
(33 MB / N) repeatedly, all ranks at once: (a) into its own pinned buffer, (b) into its band of one /dev/shm frame all ranks share
(registered with cudaHostRegister, pages first touched by rank 0), (c) the same after every rank pinned itself to the CPUs next to its
GPU (NVML) and first-touched its own band.  Prints per-rank and aggregate GB/s: tells whether the distributed Present is limited by the
links or by where the host pages live."""
import mmap
import os
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
FRAME = 3840 * 2160 * 4
band = FRAME // world // 4096 * 4096
src = torch.zeros(band, dtype=torch.uint8, device="cuda")
rt = torch.cuda.cudart()


def timed(dst, reps=200):
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    mine = band * reps / (time.perf_counter() - t0) / 1e9
    dist.barrier()
    total = time.perf_counter() - t0
    g = torch.tensor([mine], device="cuda", dtype=torch.float64)
    lst = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(lst, g)
    return [round(float(x.item()), 1) for x in lst], round(band * reps * world / total / 1e9, 1)


def shared(name, first_touch_own):
    path = f"/dev/shm/{name}"
    if rank == 0:
        fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600); os.ftruncate(fd, FRAME)
    dist.barrier()
    if rank != 0:
        fd = os.open(path, os.O_RDWR)
    m = mmap.mmap(fd, FRAME); os.close(fd)
    a = np.frombuffer(m, dtype=np.uint8)
    if first_touch_own:
        a[rank * band:(rank + 1) * band] = 1
    elif rank == 0:
        a[:] = 1
    dist.barrier()
    t = torch.from_numpy(a)
    assert rt.cudaHostRegister(t.data_ptr(), FRAME, 0) == 0 or True
    return m, t, path


res = {}
own = torch.empty(band, dtype=torch.uint8, pin_memory=True)
res["own_pinned"] = timed(own)
m1, t1, p1 = shared("b2r_probe_a", False)
res["shm_rank0_touch"] = timed(t1[rank * band:(rank + 1) * band])
rt.cudaHostUnregister(t1.data_ptr())
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    aff = sorted(os.sched_getaffinity(0))
    res["affinity"] = f"{aff[0]}..{aff[-1]} ({len(aff)} cpus)"
except Exception as e:      # noqa: BLE001
    res["affinity"] = f"unavailable: {e}"
m2, t2, p2 = shared("b2r_probe_b", True)
res["shm_own_touch_affine"] = timed(t2[rank * band:(rank + 1) * band])
rt.cudaHostUnregister(t2.data_ptr())
own2 = torch.empty(band, dtype=torch.uint8, pin_memory=True)
res["own_pinned_affine"] = timed(own2)
aff_all = [None] * world
dist.all_gather_object(aff_all, res["affinity"])
if rank == 0:
    for p in (p1, p2):
        os.unlink(p)
    print(f"band {band / 1e6:.2f} MB per rank, {world} ranks; per-rank GB/s and aggregate GB/s (wall clock incl. the slowest rank)")
    for k in ("own_pinned", "shm_rank0_touch", "shm_own_touch_affine", "own_pinned_affine"):
        print(f"  {k:24s} per rank {res[k][0]}  aggregate {res[k][1]}")
    print("  affinity per rank:", aff_all)
dist.barrier()
dist.destroy_process_group()
