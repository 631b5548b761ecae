"""Runs under the CPU emulation build of the library (B2R_LIBRARY = build/emu/libb2r_emu.so; tools/emu): the product's
kernels execute on the host, one CUDA thread per fiber.  Called by tests/test_cpu_emu.py in a subprocess, because the
library is chosen when custom_gpgpu_rasterizer_b200._lib is imported.  Prints one JSON line.

    python emu_scenarios.py parity <case>            product (emulated) vs the oracle on a golden frame case
    python emu_scenarios.py schedulers <mode>        frames in flight with a moving camera: every frame vs the oracle"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import golden_cases  # noqa: E402
import oracle_lib  # noqa: E402
import parity  # noqa: E402
from custom_gpgpu_rasterizer_b200 import _lib, scenes  # noqa: E402
from custom_gpgpu_rasterizer_b200.renderer import Renderer  # noqa: E402


def compare(r, sc, o, lists=True):
    out = {}
    fb, depth, winner, z = r.framebuffer(), r.depth(), r.winner(), r.z()
    out["winner"] = bool(np.array_equal(winner, o["winner"]))
    out["depth"] = bool(np.array_equal(depth, o["depth"]))
    out["z"] = bool(np.array_equal(parity.bits(z), parity.bits(o["z"])))
    out["colour_max_diff"] = int(np.abs(parity.channels(fb) - parity.channels(o["framebuffer"])).max())
    total, visible = r.stats()
    out["visible"] = bool(visible == o["visible_tris"] and total == o["total_tris"])
    if lists:
        counts, offsets, ids = r.bin_lists()
        want = oracle_lib.bin_sets(o["bin_counts"], o["bin_ids"])
        out["counts"] = bool(np.array_equal(counts, o["bin_counts"]))
        out["offsets"] = bool(np.array_equal(offsets, np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint32)))
        out["lists"] = bool(np.array_equal(ids, np.concatenate(want) if want else ids))       # the sets, in ascending order
        for mi, m in enumerate(sc.meshes):
            ov = r.overtex(mi, m.vertices.shape[0])
            out[f"overtex{mi}"] = bool(np.array_equal(parity.bits(ov), parity.bits(o["meshes"][mi]["overtex"])))
            out[f"visible{mi}"] = bool(np.array_equal(r.visible(mi, m.n_tris), o["meshes"][mi]["visible"]))
    return out


def main():
    what, arg = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    if what == "parity":
        sc = golden_cases.frame_cases()[arg]()
        fill_rule = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        o = oracle_lib.render(sc, fill_rule=fill_rule)
        with Renderer(sc.width, sc.height) as r:
            r.set_option(_lib.OPT_DEBUG_SURFACES, 1)
            r.set_option(_lib.OPT_FILL_RULE_TOP_LEFT, fill_rule)
            r.load_scene(sc)
            r.render_scene(sc)
            res = compare(r, sc, o)
            res["kernels"] = r.counters()["kernels_launched"]
        print(json.dumps(res))
    elif what == "direct":
        print(json.dumps(direct_path_checks(small=True, quick=(arg == "quick"))))
    elif what == "schedulers":
        import dataclasses
        sc = scenes.random_triangles_scene(160, 112, 120, seed=23, cull=2, textured=True, n_meshes=2, sample=1)
        cams = [dataclasses.replace(sc, cam_pos=(0.15 * i - 0.5, 0.05 * i, 3.0 + 0.1 * i)) for i in range(7)]
        wants = [oracle_lib.render(c, fill_rule=1) for c in cams]
        res = {"frames": []}
        with Renderer(sc.width, sc.height) as r:
            r.set_option(_lib.OPT_DEBUG_SURFACES, 1)
            if arg.startswith("ring"):
                r.set_option(_lib.OPT_FRAME_RING, int(arg[4:]))
            if arg == "serial":
                r.set_option(_lib.OPT_OVERLAP_FRONT, 0)
            if arg == "unsorted":
                r.set_option(_lib.OPT_SORT_BINS, 0)      # (the library default)
            if arg == "sorted":
                r.set_option(_lib.OPT_SORT_BINS, 1)      # lists put in order inside the frame (by the raster kernel)
                r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)     # ... with every triangle in the lists
            r.load_scene(sc)
            # all frames enqueued back to back (no synchronisation in between), Present of each into its own host buffer
            host = [np.zeros((sc.height, sc.width), np.uint32) for _ in cams]
            for i, c in enumerate(cams):
                r.render_scene(c, sync=False)
                r.present_async(host[i])
            r.present_wait()
            for i, (c, o) in enumerate(zip(cams, wants)):
                res["frames"].append(bool(np.array_equal(host[i], o["framebuffer"])))
            res["last"] = compare(r, cams[-1], wants[-1])          # the newest frame's surfaces and (canonicalised) lists
            # and frame by frame with a read-back after each
            res["stepwise"] = []
            for c, o in zip(cams[:3], wants[:3]):
                r.render_scene(c, sync=False)
                res["stepwise"].append(compare(r, c, o, lists=False))
        print(json.dumps(res))


def direct_path_checks(small=True, quick=False):
    """Direct path of the setup kernel (B2R_OPT_DIRECT_MAX_AREA) against the all-binned path on the same frames: every
    surface, the statistics and the (complete) bins must be identical, for several bin grids, thresholds, fill rules,
    sort-first tile ownership and with frames in flight (the key buffer must be empty again after every frame).
    Also used on the GPU (tests/test_gpu_parity.py) with larger frames."""
    import dataclasses
    w, h = (160, 112) if small else (640, 360)
    n = 1 if small else 6
    cases = {
        "tiny": scenes.random_triangles_scene(w, h, 900 * n, seed=71, cull=2, size=0.05),
        "mixed": scenes.random_triangles_scene(w, h, 500 * n, seed=72, cull=0, size=0.3, textured=True, n_meshes=2, sample=1),
        "spheres": scenes.sphere_grid_scene(w, h, 3, 2, 21 if small else 41, 24 if small else 60),
        "hostile": scenes.hostile_scene(w, h, 300 * n, seed=73, textured=True),
        "sliver": scenes.sliver_scene(w, h, 400 * n, seed=74),
    }
    res = {}

    def surfaces(r):
        return [r.framebuffer().copy(), r.depth().copy(), r.winner().copy(), parity.bits(r.z()).copy()]

    for name, sc in cases.items():
        if quick and name in ("spheres", "sliver"):
            continue
        for grid in ("default", "32x32", "24x20", "40x8"):
            if grid != "default" and (name not in ("tiny", "mixed") or (quick and grid != "24x20")):
                continue
            for fill in (1, 0):
                if fill == 0 and (grid != "default" or (quick and name != "mixed")):
                    continue
                with Renderer(sc.width, sc.height) as r:
                    r.set_option(_lib.OPT_DEBUG_SURFACES, 1)
                    r.set_option(_lib.OPT_COUNT_FRAGMENTS, 1)
                    r.set_option(_lib.OPT_FILL_RULE_TOP_LEFT, fill)
                    r.load_scene(sc)
                    if grid != "default":
                        bw, bh = (int(v) for v in grid.split("x"))
                        r.setup_bins((sc.width + bw - 1) // bw, (sc.height + bh - 1) // bh, bw, bh)
                    r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)
                    r.render_scene(sc)
                    want, wc, wstats = surfaces(r), r.counters(), r.stats()
                    wbins = r.bin_lists()
                    for area in ((16, 4096) if quick else (1, 16, 64, 4096)):
                        r.set_option(_lib.OPT_DIRECT_MAX_AREA, area)
                        r.render_scene(sc)
                        got, gc = surfaces(r), r.counters()
                        ok = all(np.array_equal(a, b) for a, b in zip(want, got))
                        ok = ok and gc["pairs"] == wc["pairs"] and gc["fragments"] == wc["fragments"] and r.stats() == wstats
                        r.render_scene(sc)      # again: the key buffer was left empty
                        ok = ok and all(np.array_equal(a, b) for a, b in zip(want, surfaces(r)))
                        gbins = r.bin_lists()      # complete bins (re-run without the path), then the path again
                        ok = ok and all(np.array_equal(a, b) for a, b in zip(wbins, gbins))
                        r.render_scene(sc)
                        ok = ok and all(np.array_equal(a, b) for a, b in zip(want, surfaces(r)))
                        res[f"{name}/{grid}/fill{fill}/area{area}"] = bool(ok)
    # sort-first with a replicated front end: each owner renders its tiles only; composed = the whole frame
    sc = cases["mixed"]
    with Renderer(sc.width, sc.height) as r:
        r.load_scene(sc)
        r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)
        r.render_scene(sc)
        want = r.framebuffer().copy()
        r.set_option(_lib.OPT_DIRECT_MAX_AREA, 64)
        composed = np.zeros_like(want)
        tiles_x = (sc.width + 15) // 16
        ty, tx = np.meshgrid(np.arange(sc.height) // 16, np.arange(sc.width) // 16, indexing="ij")
        for rank in range(3):
            r.set_option(_lib.OPT_TILE_OWNER_COUNT, 3)
            r.set_option(_lib.OPT_TILE_OWNER_RANK, rank)
            r.render_scene(sc)
            fb = r.framebuffer()
            mine = (tx + ty * tiles_x) % 3 == rank
            composed[mine] = fb[mine]
        res["owners3"] = bool(np.array_equal(composed, want))
    # frames in flight with a moving camera: three-stream pipeline and the frame ring
    sc = cases["tiny"]
    cams = [dataclasses.replace(sc, cam_pos=(0.15 * i - 0.5, 0.05 * i, 3.0 + 0.1 * i)) for i in range(6)]
    for mode in ("pipeline", "ring3"):
        with Renderer(sc.width, sc.height) as r:
            if mode == "ring3":
                r.set_option(_lib.OPT_FRAME_RING, 3)
            r.load_scene(sc)
            r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)
            wants = []
            for c in cams:
                r.render_scene(c)
                wants.append(r.framebuffer().copy())
            r.set_option(_lib.OPT_DIRECT_MAX_AREA, 64)
            host = [np.zeros((sc.height, sc.width), np.uint32) for _ in cams]
            for i, c in enumerate(cams):
                r.render_scene(c, sync=False)
                r.present_async(host[i])
            r.present_wait()
            res[f"inflight/{mode}"] = bool(all(np.array_equal(a, b) for a, b in zip(host, wants)))
    return res


def sharded(world, mode, n_meshes=2, n_tris=700):
    """N contexts of this process as the N ranks of the sharded sort-first front end: every rank renders the same frames
    (moving camera, all enqueued back to back), rank 0 presents each composed frame; compared with the oracle's frames."""
    import ctypes as C
    import dataclasses
    lib = _lib.lib
    if os.environ.get("B2R_EMU_SCENE") == "hostile":      # triangles through the camera plane, w = 0, NaN / inf / huge coordinates, degenerate ones
        sc = scenes.hostile_scene(160, 112, 400, seed=43, textured=True)
    elif os.environ.get("B2R_EMU_SCENE") == "tiny":       # mostly triangles of a few pixels (the direct path: key-buffer blocks pushed to the owners) among larger ones
        sc = scenes.random_triangles_scene(160, 112, n_tris * 3, seed=31, cull=2, textured=True, n_meshes=n_meshes, sample=1, size=0.12)
    else:
        sc = scenes.random_triangles_scene(160, 112, n_tris, seed=29, cull=2, textured=True, n_meshes=n_meshes, sample=1)
    cams = [dataclasses.replace(sc, cam_pos=(0.15 * i - 0.5, 0.05 * i, 3.0 + 0.1 * i)) for i in range(7)]
    wants = [oracle_lib.render(c, fill_rule=1) for c in cams]
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
    host = [np.zeros((sc.height, sc.width), np.uint32) for _ in cams]
    for i, c in enumerate(cams):
        for r in ranks:
            r.render_scene(c, sync=False)
        ranks[0].present_async(host[i])
    ranks[0].present_wait()
    res = {"frames": [bool(np.array_equal(h, o["framebuffer"])) for h, o in zip(host, wants)]}
    res["last_framebuffer"] = bool(np.array_equal(ranks[0].framebuffer(), wants[-1]["framebuffer"]))
    res["visible_direct"] = int(sum(r.stats()[1] for r in ranks)) == int(wants[-1]["visible_tris"])
    res["pairs_direct"] = int(sum(r.counters()["pairs"] for r in ranks)) == int(wants[-1]["bin_counts"].sum())
    # the ranks' lists of the newest frame concatenate, in rank order, to the single-GPU lists (= the oracle's sets, ascending).
    # Complete lists need the direct path off (it hands small triangles to the owners as key-buffer blocks, not as list
    # entries): the newest frame once more, every triangle binned, on every rank
    direct = ranks[0].get_option(_lib.OPT_DIRECT_MAX_AREA)
    for r in ranks:
        r.set_option(_lib.OPT_DIRECT_MAX_AREA, 0)
    for r in ranks:
        r.render_scene(cams[-1], sync=False)
    res["last_framebuffer_binned"] = bool(np.array_equal(ranks[0].framebuffer(), wants[-1]["framebuffer"]))
    per_rank = [r.bin_lists() for r in ranks]
    for r in ranks:
        r.set_option(_lib.OPT_DIRECT_MAX_AREA, direct)
    n_bins = len(per_rank[0][0])
    cat = []
    for b in range(n_bins):
        for counts, offsets, ids in per_rank:
            cat.append(ids[offsets[b]:offsets[b] + counts[b]])
    want = oracle_lib.bin_sets(wants[-1]["bin_counts"], wants[-1]["bin_ids"])
    res["lists"] = bool(np.array_equal(np.concatenate(cat), np.concatenate(want)))
    res["visible"] = int(sum(r.stats()[1] for r in ranks)) == int(wants[-1]["visible_tris"])
    # distributed Present: every rank keeps its tiles' colours and copies them into ONE host frame all ranks share
    # (b2r_present_tiles_async); two frames in flight into the two halves of the host buffer
    shared = np.zeros((2, sc.height, sc.width), np.uint32)
    for r in ranks:
        r.sync()
    for r in ranks:
        r.register_host_frame(shared)
        r.set_option(_lib.OPT_SHARD_LOCAL_COLOUR, 1)
    for i in (2, 5):
        for r in ranks:
            r.render_scene(cams[i], sync=False)
            r.present_tiles_async((i & 1) * sc.width * sc.height * 4)
    for r in ranks:
        r.present_wait()
    res["distributed_present"] = bool(np.array_equal(shared[0], wants[2]["framebuffer"]) and np.array_equal(shared[1], wants[5]["framebuffer"]))
    for r in ranks:
        r.set_option(_lib.OPT_SHARD_LOCAL_COLOUR, 0)
        r.register_host_frame(None)
    # ... and by row bands out of the frame composed on rank 0 (b2r_present_band_async), again two frames in flight
    shared[:] = 0
    for r in ranks:
        r.register_host_frame(shared)
    for i in (4, 1):
        for r in ranks:
            r.render_scene(cams[i], sync=False)
            r.present_band_async((i & 1) * sc.width * sc.height * 4)
    for r in ranks:
        r.present_wait()
    res["band_present"] = bool(np.array_equal(shared[0], wants[4]["framebuffer"]) and np.array_equal(shared[1], wants[1]["framebuffer"]))
    for r in ranks:
        r.register_host_frame(None)
    for r in ranks:
        r.render_scene(cams[3], sync=False)
    res["composed_on_rank0_again"] = bool(np.array_equal(ranks[0].framebuffer(), wants[3]["framebuffer"]))
    for r in ranks:
        r.sync()
    for r in ranks:
        r._check(lib.b2r_shard_release(r.handle))
        r.close()
    print(json.dumps(res))


def dynamic(mode):
    """Scene changes between frames in flight: a mesh's world matrix (b2r_mesh_set_world), then a third mesh - every frame,
    enqueued without synchronisation and presented into its own buffer, against the oracle's frame of the scene as it was
    when the frame was issued."""
    import dataclasses
    from custom_gpgpu_rasterizer_b200.renderer import host_mesh_world
    from custom_gpgpu_rasterizer_b200.scenefile import MeshData
    lib = _lib.lib
    base = scenes.random_triangles_scene(160, 112, 150, seed=37, cull=2, textured=True, n_meshes=2)
    extra = scenes.random_triangles_scene(160, 112, 60, seed=38, cull=2).meshes[0]

    def moved(sc, pos):
        m0 = sc.meshes[0]
        return dataclasses.replace(sc, meshes=[MeshData(m0.vertices, m0.indices, pos, m0.tex_ids)] + list(sc.meshes[1:]))
    s1 = moved(base, (0.4, -0.2, 0.3))
    s2 = dataclasses.replace(s1, meshes=list(s1.meshes) + [MeshData(extra.vertices, extra.indices, (0.0, 0.3, 0.5), (-1, -1, -1, -1))])
    plan = [base, base, s1, s1, s1, s2, s2, s2]
    wants = {id(x): oracle_lib.render(x, fill_rule=1) for x in (base, s1, s2)}
    host = [np.zeros((base.height, base.width), np.uint32) for _ in plan]
    with Renderer(base.width, base.height) as r:
        if mode.startswith("ring"):
            r.set_option(_lib.OPT_FRAME_RING, int(mode[4:]))
        r.load_scene(base)
        cur = base
        for i, sc in enumerate(plan):
            if sc is s1 and cur is base:
                w = host_mesh_world(s1.meshes[0].position)
                r._check(lib.b2r_mesh_set_world(r.handle, 0, w.ctypes.data_as(_lib._fp)))
            if sc is s2 and cur is s1:
                m = s2.meshes[2]
                r.add_mesh(m.vertices, m.indices, host_mesh_world(m.position), m.tex_ids)
            cur = sc
            r.render_scene(sc, sync=False)
            r.present_async(host[i])
        r.present_wait()
    print(json.dumps({"frames": [bool(np.array_equal(h, wants[id(sc)]["framebuffer"])) for h, sc in zip(host, plan)]}))


def replicated(world):
    """north_star's literal sort-first form: every rank bins the full mesh, rasterises only the tiles it owns (tile index mod
    ranks) into a packed tile-major buffer, and rank 0 scatters the peers' buffers into its framebuffer (b2r_unpack_tiles) -
    what multi_gpu.SortFirstGather does with an NCCL gather in between.  N contexts of this process as the N ranks."""
    import ctypes as C
    from custom_gpgpu_rasterizer_b200.multi_gpu import packed_pixels
    lib = _lib.lib
    sc = scenes.random_triangles_scene(333, 217, 400, seed=31, cull=2, textured=True, n_meshes=2)       # ragged resolution: partial tiles
    o = oracle_lib.render(sc, fill_rule=1)
    ranks = [Renderer(sc.width, sc.height, device=d) for d in range(world)]
    packed = [None]
    for k, r in enumerate(ranks):
        r.set_option(_lib.OPT_TILE_OWNER_COUNT, world)
        r.set_option(_lib.OPT_TILE_OWNER_RANK, k)
        r.load_scene(sc)
        if k:
            n = C.c_uint64(0)
            r._check(lib.b2r_packed_size(r.handle, k, world, C.byref(n)))
            assert n.value == packed_pixels(sc.bins_x * sc.bins_y, k, world)
            packed.append(np.zeros(int(n.value), np.uint32))
            r.set_option(_lib.OPT_PACKED_OUTPUT, 1)
            r.bind_framebuffer(packed[k].ctypes.data)
    for r in ranks:
        r.render_scene(sc)
    for k in range(1, world):
        ranks[0]._check(lib.b2r_unpack_tiles(ranks[0].handle, C.c_void_p(packed[k].ctypes.data), k, world, None))
    res = {"framebuffer": bool(np.array_equal(ranks[0].framebuffer(), o["framebuffer"])),
           "pairs": int(sum(r.counters()["pairs"] for r in ranks)) == int(o["bin_counts"].sum())}
    for r in ranks:
        r.close()
    print(json.dumps(res))


if __name__ == "__main__":
    if sys.argv[1] == "dynamic":
        dynamic(sys.argv[2])
        sys.exit(0)
    if sys.argv[1] == "replicated":
        replicated(int(sys.argv[2]))
        sys.exit(0)
    if sys.argv[1] == "sharded":
        sharded(int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 2, int(sys.argv[5]) if len(sys.argv) > 5 else 700)
        sys.exit(0)
    main()
