"""CPU (gloo, world_size 2) tests of the sort-first host logic: tile ownership, the packed tile-major
layout (numpy models of B2R_OPT_PACKED_OUTPUT / b2r_unpack_tiles) and the gather plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from custom_gpgpu_rasterizer_b200 import multi_gpu as mg  # noqa: E402


def test_ownership_partitions_all_bins():
    for n_bins, world in [(32400, 8), (8160, 3), (7, 4), (1, 2)]:
        seen = np.concatenate([mg.owned_bins(n_bins, r, world) for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(n_bins))
        sizes = [len(mg.owned_bins(n_bins, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("w,h,bw,bh", [(64, 48, 16, 16), (70, 37, 16, 16), (160, 96, 80, 48)])
def test_pack_unpack_roundtrip(w, h, bw, bh):
    bins_x, bins_y = (w + bw - 1) // bw, (h + bh - 1) // bh
    rng = np.random.default_rng(1)
    img = rng.integers(0, 1 << 32, (h, w), dtype=np.uint64).astype(np.uint32)
    world = 3
    out = np.zeros_like(img)
    for r in range(world):
        packed = mg.pack_reference(img, r, world, bins_x, bins_y, bw, bh)
        sub = ((bw + 15) // 16) * ((bh + 15) // 16)
        assert packed.size == mg.packed_pixels(bins_x * bins_y, r, world, sub)
        mg.unpack_reference(packed, out, r, world, bins_x, bins_y, bw, bh)
    assert np.array_equal(out, img)


def _worker(rank, world, port, w, h, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bins_x, bins_y = (w + 15) // 16, (h + 15) // 16
        rng = np.random.default_rng(7)
        full = rng.integers(0, 1 << 31, (h, w), dtype=np.int64).astype(np.int32)   # the frame every rank would agree on
        n_bins = bins_x * bins_y
        max_px = max(mg.packed_pixels(n_bins, k, world) for k in range(world))
        local = np.zeros(max_px, np.int32)
        mine = mg.pack_reference(full, rank, world, bins_x, bins_y)
        local[: mine.size] = mine
        bufs = mg.gather_packed(torch.from_numpy(local), rank, world, dst=0)
        if rank == 0:
            out = np.zeros_like(full)
            for k in range(world):
                mg.unpack_reference(bufs[k].numpy(), out, k, world, bins_x, bins_y)
            q.put(bool(np.array_equal(out, full)))
    finally:
        dist.destroy_process_group()


def test_gloo_gather_composes_full_frame():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, 100, 70, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_prim_slices_partition_the_primitives():
    for n in (0, 1, 255, 256, 257, 10_000_000, 1_234_567):
        for world in (2, 3, 4, 8):
            cuts = [mg.prim_slice(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            assert all(lo % 256 == 0 for lo, _ in cuts)


def _bin_slice(rects, lo, hi, bins_x, n_bins):
    """numpy model of one rank's front end: ascending per-tile lists of the primitives [lo, hi)."""
    lists = [[] for _ in range(n_bins)]
    for prim in range(lo, hi):
        x0, x1, y0, y1 = rects[prim]
        for ty in range(y0, y1 + 1):
            for tx in range(x0, x1 + 1):
                lists[tx + ty * bins_x].append(prim)
    return lists


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bins_x, bins_y, n = 12, 9, 3000
        rng = np.random.default_rng(11)
        x0 = rng.integers(0, bins_x, n); y0 = rng.integers(0, bins_y, n)
        rects = np.stack([x0, np.minimum(x0 + rng.integers(0, 3, n), bins_x - 1), y0, np.minimum(y0 + rng.integers(0, 3, n), bins_y - 1)], 1)
        lo, hi = mg.prim_slice(n, rank, world)
        mine = _bin_slice(rects, lo, hi, bins_x, bins_x * bins_y)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        # the owner of a tile walks the ranks' lists in rank order: must be the single-GPU ascending list
        single = _bin_slice(rects, 0, n, bins_x, bins_x * bins_y)
        ok = True
        for t in mg.owned_bins(bins_x * bins_y, rank, world):
            cat = [p for k in range(world) for p in everyone[k][t]]
            ok &= cat == single[t]
        q.put(ok)
    finally:
        dist.destroy_process_group()


def test_gloo_sharded_front_end_lists_concatenate_in_rank_order():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True and q.get(timeout=10) is True


# ---- frame ring of the sharded front end: ordering rules under random timing (tests/ring_protocol_model.py)
@pytest.mark.parametrize("ranks,depth", [(2, 2), (2, 3), (4, 3), (8, 3), (8, 4)])
def test_frame_ring_rules_order_every_list_set_access(ranks, depth):
    import ring_protocol_model as model
    for seed in range(40):
        bad, _ = model.simulate(ranks, depth, 24, np.random.default_rng(seed))
        assert not bad, f"seed {seed}: {bad[:4]}"


def test_frame_ring_model_has_teeth():
    """Each rule is needed: without it the model finds an unordered access for some timing."""
    import ring_protocol_model as model
    found = {"no_done_wait": 0, "epoch_off_by_one": 0, "unchained_rasters": 0}
    for seed in range(40):
        rng = lambda: np.random.default_rng(seed)
        found["no_done_wait"] += any(v[0] == "S1" for v in model.simulate(8, 3, 24, rng(), done_wait=False)[0])
        found["epoch_off_by_one"] += any(v[0] == "S1" for v in model.simulate(8, 3, 24, rng(), done_lag=1)[0])
        found["unchained_rasters"] += any(v[0] == "S3" for v in model.simulate(8, 3, 24, rng(), chain_rasters=False)[0])
    assert all(n > 0 for n in found.values()), found


def test_frame_ring_overlaps_front_ends():
    """What the ring is for: with D frames in flight the steady-state frame time approaches the longest single stage
    instead of the sum of the stages of one frame."""
    import ring_protocol_model as model
    serial = np.diff(model.simulate(8, 1, 40, np.random.default_rng(3), skew=0.0)[1])[10:].mean()
    ring = np.diff(model.simulate(8, 3, 40, np.random.default_rng(3), skew=0.0)[1])[10:].mean()
    assert ring < 0.75 * serial, (ring, serial)


def test_shard_modes_must_agree():
    mg.check_shard_modes([(3, 1), (3, 1)])
    with pytest.raises(RuntimeError):
        mg.check_shard_modes([(3, 1), (0, 1)])
