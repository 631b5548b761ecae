#!/usr/bin/env python
"""Headline benchmark of the B200 rendering path (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload spheres10m|fill64]

One "step" = one frame of the workload through the per-frame hot path.  At N=1 the workload is the
configuration BASELINE.json's metric is quoted on: the 10 000 000-triangle sphere grid at 3840x2160
(config 3).  N>1 launches one rank per GPU (torchrun); the frame is sharded sort-first (each rank owns an
interleaved set of 16x16 tiles, transforms and bins the full mesh, rasterises only its tiles, and the
framebuffer is gathered to rank 0 over NCCL/NVLink).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: this image exports NCCL_DEBUG=VERSION, which makes NCCL print its version
# banner to stdout (so does WARN); drop those levels, and send an explicitly requested INFO/TRACE log to stderr
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    del os.environ["NCCL_DEBUG"]
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from custom_gpgpu_rasterizer_b200 import scenes  # noqa: E402  (numpy only)

WIDTH, HEIGHT = 3840, 2160
METRIC = "frames/s at 4K (device-timed)"      # BASELINE.json's metric; the same string on both arms


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_scene(workload: str):
    if workload == "spheres10m":
        return scenes.sphere_grid_scene(WIDTH, HEIGHT)                      # BASELINE config 3
    if workload == "spheres_small":                                        # quick functional check
        return scenes.sphere_grid_scene(WIDTH, HEIGHT, nx=5, ny=5)
    if workload == "fill64":
        return make_fill_scene()                                            # BASELINE config 4
    if workload in ("vehicle480", "vehicle1080"):
        return make_vehicle_scene(*((640, 480) if workload == "vehicle480" else (1920, 1080)))   # BASELINE configs 1 / 2
    raise SystemExit(f"unknown workload {workload}")


def make_fill_scene():
    """BASELINE config 4: 64 stacked full-screen quads at 4K, linear sampling.  The diffuse map is the reference's bundled
    uv_grid_2.png when oracle/build_ref.sh has copied the assets next to the reference binary - a data
    file, decoded by the product's own PNG decoder - else a synthetic 1024^2 checker; `fill_texture` says which."""
    path = os.path.join(ROOT, "oracle", "_ref", "resources", "uv_grid_2.png")
    tex, src = None, "synthetic 1024x1024 checker"
    if os.path.exists(path):
        try:
            from custom_gpgpu_rasterizer_b200.renderer import decode_png
            tex = decode_png(path)
            src = f"uv_grid_2.png ({tex.shape[1]}x{tex.shape[0]}, the reference's bundled map)"
        except Exception:      # noqa: BLE001
            tex = None
    sc = scenes.fill_quads_scene(WIDTH, HEIGHT, layers=64, sample=1, texture=tex)
    sc.fill_texture = src
    return sc


def make_vehicle_scene(width, height):
    """BASELINE configs 1 and 2: the reference's bundled vehicle mesh (the REFERENCE parser's output, committed as a golden
    vector), camera {0,5,65} / fov 45 (main.cpp:421), back-face culling, point sampling, the four texture slots wired as in
    main.cpp:172-175.  The maps are the reference's own PNGs when oracle/build_ref.sh has copied them next to the reference
    binary (data files, decoded by the product's PNG decoder), else synthetic 1024^2 maps - `data` in the JSON line says which."""
    from custom_gpgpu_rasterizer_b200.renderer import decode_png
    from custom_gpgpu_rasterizer_b200.scenefile import MeshData, Scene
    g = np.load(os.path.join(ROOT, "tests", "golden", "objparser_reference.npz"))
    res = os.path.join(ROOT, "oracle", "_ref", "resources")
    names = [os.path.join(res, f"vehicle_{k}.png") for k in ("diffuse", "normal", "specular", "gloss")]
    if all(os.path.exists(n) for n in names):
        tex, src = [decode_png(n) for n in names], "bundled"
    else:
        tex, src = [scenes.checker_texture(1024, 1), scenes.normal_texture(1024, 2), scenes.noise_texture(1024, 3), scenes.noise_texture(1024, 4)], "synthetic"
    sc = Scene(width, height, meshes=[MeshData(g["res_vehicle_v"], g["res_vehicle_i"], (0.0, 0.0, 0.0), (0, 1, 2, 3))], textures=tex,
               cull=0, sample=0, vis=0, cam_pos=(0.0, 5.0, 65.0), fov_deg=45.0, near=0.1, far=100.0, name=f"vehicle_{width}x{height}")
    sc.texture_source = src
    return sc


# ------------------------------------------------------------------------------------ BASELINE config 5 (optional workload)
def make_views_scene():
    """172 instances of the reference's bundled vehicle mesh (the REFERENCE parser's output, committed as a golden
    vector in tests/golden/objparser_reference.npz; textures are synthetic 1024^2 maps) on a 14x13 lattice, and the
    1024 cameras of SURVEY section 8d: Camera(pos, 45), Pitch(-19.654), Yaw(360 i / 1024), pos = 743.3 * forward."""
    from custom_gpgpu_rasterizer_b200.renderer import host_camera
    from custom_gpgpu_rasterizer_b200.scenefile import MeshData, Scene
    g = np.load(os.path.join(ROOT, "tests", "golden", "objparser_reference.npz"))
    v, i = g["res_vehicle_v"], g["res_vehicle_i"]
    tex = [scenes.checker_texture(1024, 1), scenes.normal_texture(1024, 2), scenes.noise_texture(1024, 3), scenes.noise_texture(1024, 4)]
    meshes = [MeshData(v, i, ((k % 14 - 6.5) * 45.0, 0.0, (k // 14 - 6.0) * 45.0), (0, 1, 2, 3)) for k in range(172)]
    sc = Scene(1920, 1080, meshes=meshes, textures=tex, cull=0, sample=0, vis=0, fov_deg=45.0, near=1.0, far=2000.0, name="vehicles172_views1024")
    cams = []
    for k in range(1024):
        pitch, yaw = -19.654, 360.0 * k / 1024
        fwd = host_camera((0.0, 0.0, 0.0), 45.0, sc.width, sc.height, 1.0, 2000.0, pitch, yaw)["cam_fwd"]
        pos = tuple(float(np.float32(743.3) * f) for f in fwd)
        cams.append(host_camera(pos, 45.0, sc.width, sc.height, 1.0, 2000.0, pitch, yaw))
    return sc, cams


def run_views(args):
    """View-parallel batch (default): view i is rendered by rank i mod N, no communication; K views per rank are timed.
    --views-mode sortfirst (N > 1): every view is rendered by all N ranks together through the sharded sort-first front end
    (BASELINE config 5's "plus sort-first per view"); K views in total are timed, the composed frames live on rank 0."""
    import torch
    import torch.distributed as dist
    from custom_gpgpu_rasterizer_b200.renderer import Renderer
    refuse_emulation()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sc, cams = make_views_scene()
    r = Renderer(sc.width, sc.height, device=local)
    r.load_scene(sc)
    stream = torch.cuda.Stream(device=dev)
    r.set_stream(stream.cuda_stream)
    sort_first = args.views_mode == "sortfirst" and world > 1
    shard = None
    if sort_first:
        from custom_gpgpu_rasterizer_b200.multi_gpu import ShardedSortFirst
        shard = ShardedSortFirst(r, rank, world)
        mine = [cams[k % len(cams)] for k in range(max(args.warmup, 3) + args.steps)]        # the same views on every rank
    else:
        mine = [cams[(rank + k * world) % len(cams)] for k in range(max(args.warmup, 3) + args.steps)]
    views_per_step = 1 if sort_first else world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for cam in mine[: max(args.warmup, 3)]:
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
    r.sync()
    timed = mine[max(args.warmup, 3):]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for cam in timed:
                r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
            e1.record(stream)
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = r.counters()["kernels_launched"]
    pinned = [torch.empty((sc.height, sc.width), dtype=torch.int32, pin_memory=True) for _ in range(2)]
    host_fb = [p.numpy().view(np.uint32) for p in pinned]
    barrier()
    t0 = time.perf_counter()
    for k, cam in enumerate(timed):
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
        if not sort_first or rank == 0:        # sort-first: the composed frame is rank 0's
            r.present_async(host_fb[k & 1])
    if not sort_first or rank == 0:
        r.present_wait()
    else:
        r.sync()
    barrier()
    e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    total, visible = r.stats()
    if sort_first:
        total = sc.n_tris                       # a rank's own counters cover its primitive slice only
    if shard is not None:
        r.sync()
        barrier()
        shard.close()
    r.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    t = float(ms.item()) / 1e3
    world_views = views_per_step
    print(json.dumps({
        "metric": "views/s (BASELINE config 5, device-timed)", "value": world_views * args.steps / t, "unit": "views/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if sort_first else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE config 5: 172 instances of the bundled vehicle mesh (2 001 736 triangles, one vertex buffer + 172 world matrices), "
                               "1920x1080, 1024 cameras on a ring, PBR shading with four 1024^2 maps",
                   "triangles": total, "visible_triangles_last_view": visible, "resolution": "1920x1080"},
        "parallelism": (f"sort-first per view, {world} ranks (sharded front end, frames composed on rank 0)" if sort_first else f"view-parallel, {world} rank(s): view i -> rank i mod N, no communication"),
        "mtris_per_s": total * world_views * args.steps / t / 1e6,
        "e2e": {"value": world_views * args.steps / float(e2e.item()), "unit": "views/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": sc.width * sc.height * 4},
        "gpu_launches": launches * args.steps, "clocks": clocks.summary()}), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, n in enumerate(names):
                if len(s) > 3 + i and s[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def calibrate_bands(torch, dist, dev, shared, scene, rank, world, barrier, reps=30):
    """Rows [begin, end) of the frame per rank for the distributed Present, in proportion to the device->host rate every rank
    reaches while ALL ranks copy into the shared host frame (whole 16-row tile rows; every rank gets at least one)."""
    rows16 = (scene.height + 15) // 16
    n = scene.width * (scene.height // world // 16 * 16 or 16)               # pixels of an equal band
    src = torch.zeros(n, dtype=torch.int32, device=dev)
    dst = torch.from_numpy(shared[:scene.width * scene.height].view(np.int32))[rank * n: rank * n + n] if rank * n + n <= scene.width * scene.height else None
    if dst is None:
        dst = torch.from_numpy(shared[:n].view(np.int32))
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    rate = n * 4 * reps / (time.perf_counter() - t0)
    barrier()
    rates = [None] * world
    dist.all_gather_object(rates, rate)
    w = np.array([max(float(x), 1.0) for x in rates])
    share = np.maximum(np.round(np.cumsum(w) / w.sum() * rows16).astype(int), np.arange(1, world + 1))     # cumulative tile rows, >= 1 each
    share = np.minimum(share, rows16 - (world - 1 - np.arange(world)))
    share = np.maximum.accumulate(np.clip(share, 0, rows16))      # (fewer tile rows than ranks: some bands stay empty)
    share[-1] = rows16
    bounds, lo = [], 0
    for hi in share:
        bounds.append((min(lo * 16, scene.height), min(int(hi) * 16, scene.height)))
        lo = int(hi)
    return bounds


class SharedHostFrames:
    """Host memory shared by the ranks of one box (a file in /dev/shm mapped by every rank): the frame a distributed Present
    composes.  Rank 0 creates and removes the file."""

    def __init__(self, name: str, nbytes: int, rank: int, barrier):
        import mmap
        self.path, self.rank = os.path.join("/dev/shm", name), rank
        if rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, nbytes)
        barrier()
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self.map = mmap.mmap(fd, nbytes)
        os.close(fd)
        self.array = np.frombuffer(self.map, dtype=np.uint32)
        barrier()

    def close(self):
        self.array = None
        try:
            self.map.close()
        except BufferError:
            pass
        if self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)


# ------------------------------------------------------------------------------------ CPU arm
ERE_REF = os.path.join(ROOT, "oracle", "_ref", "libere_ref.so")
ERE_PORT = os.path.join(ROOT, "oracle", "liborc_erenderer.so")


def cpu_reference_arm(scene, steps: int, warmup: int):
    """The reference's CPU software path, one thread (it is single-threaded): Elite::Renderer::Render, steady_clock around
    the call, scene load excluded.  Runs the reference's OWN renderer when it was built here (oracle/_ref/libere_ref.so =
    CustomCUDARasterizer/ERenderer.cpp compiled unmodified by oracle/build_ref_erenderer.sh; kind "reference"), else the
    restatement oracle/liborc_erenderer.so, which tests/test_cpu_erenderer.py pins against it bit for bit (kind "port").
    Returns (mean seconds per frame, per-frame seconds, kind)."""
    kind = "reference" if os.path.exists(ERE_REF) and not os.environ.get("B2R_CPU_BASELINE_PORT") else "port"
    if kind == "port" and not os.path.exists(ERE_PORT):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc_erenderer.so"])
    L = C.CDLL(ERE_REF if kind == "reference" else ERE_PORT)
    L.ere_load.restype = C.c_void_p
    L.ere_load.argtypes = [C.c_char_p, C.c_uint64]
    L.ere_render.restype = C.c_double
    L.ere_render.argtypes = [C.c_void_p] + [C.c_uint32] * 4
    L.ere_free.argtypes = [C.c_void_p]
    blob = scene.to_blob()
    S = L.ere_load(blob, len(blob))
    times = []
    for i in range(warmup + steps):
        t = L.ere_render(S, 0, 0xFFFFFFFF, 0, 0xFFFFFFFF)
        if t < 0:
            raise RuntimeError("the CPU renderer rejected the scene")
        if i >= warmup:
            times.append(t)
    L.ere_free(S)
    return float(np.mean(times)), times, kind


CPU_KIND_NOTE = {"reference": "the reference's own Elite::Renderer (CustomCUDARasterizer/ERenderer.cpp compiled unmodified, oracle/build_ref_erenderer.sh)",
                 "port": "ERenderer restatement (oracle/erenderer_oracle.cpp; bit-identical to the reference's renderer on tests/test_cpu_erenderer.py's scenes)"}


def reference_cuda_arm(scene, frames: int, timeout_s: int):
    """The reference's own CUDA renderer (oracle/_ref/ref_render = the unmodified reference sources compiled for sm_100a, headless),
    timed on GPU 0 of this box on the same scene: wall clock of its RenderAuto, which synchronises the device inside every frame.
    A reported baseline next to the CPU one; returns None when the binary is absent or the run fails / exceeds its time-out."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_render")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_render not built"}
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "scene.b2rs")
        with open(sp, "wb") as f:
            f.write(scene.to_blob())
        try:
            res = subprocess.run([exe, "time", sp, str(frames)], capture_output=True, text=True, timeout=timeout_s)
        except subprocess.TimeoutExpired:
            return {"unavailable": f"ref_render did not finish 1 + {frames} frames within {timeout_s} s"}
    if res.returncode != 0:
        return {"unavailable": f"ref_render exit code {res.returncode}: {res.stderr.strip()[-200:]}"}
    try:
        out = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        return {"unavailable": "ref_render printed no timing line"}
    return {"value": 1e3 / out["ms_per_frame"], "unit": "frames/s", "ms_per_frame": out["ms_per_frame"], "frames": out["frames"],
            "visible_triangles": out["visible_triangles"], "kind": "reference",
            "sample": f"the full frame of this workload, {out['frames']} frame(s) after 1 warm-up, through the reference's CUDARenderer::RenderAuto "
                      "(unmodified sources built by oracle/build_ref.sh), wall clock; 256-triangle batches with a device synchronisation after every launch"}


def run_reference(args, scene):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the driver's --steps / --warmup are honoured: the full frame costs ~0.55 s on one core of the GPU box, so the default
    # 20 + 5 frames take ~15 s; only an extravagant request is cut down to stay within minutes (and the line says so)
    steps, warmup = min(args.steps, 200), min(args.warmup, 50)
    mean_s, _, kind = cpu_reference_arm(scene, steps, warmup)
    fps = 1.0 / mean_s
    line = {
        "impl": "reference", "metric": METRIC, "timing": "host steady_clock around Render (this arm runs on the host cores)",
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": mean_s * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_of(scene),
        "mtris_per_s": scene.n_tris / mean_s / 1e6,
        "config": workload_config(args.workload, scene),
        "parallelism": "the reference's CPU renderer, one host thread (it is single-threaded)",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind,
                         "sample": f"full frame of the workload, {steps} frames after {warmup} warm-up; {CPU_KIND_NOTE[kind]}; single-threaded like the reference",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.ref_cuda:
        line["reference_cuda_path"] = reference_cuda_arm(scene, args.ref_cuda, args.ref_cuda_timeout)
    print(json.dumps(line), flush=True)


def data_of(scene):
    src = getattr(scene, "texture_source", None)
    return "synthetic" if src is None else f"the reference's bundled vehicle mesh, {src} texture maps"


def workload_config(name, scene):
    return {"workload": {"spheres10m": "BASELINE config 3: 25x25 UV-sphere grid, 10 000 000 triangles / 5 176 250 vertices, 3840x2160, 16x16 tiles (240x135), back-face culling, vertex colours",
                         "fill64": "BASELINE config 4: 64 stacked full-screen textured quads, 3840x2160, linear sampling, no culling",
                         "spheres_small": "5x5 spheres (400 000 triangles), 3840x2160",
                         "vehicle480": "BASELINE config 1: the reference's bundled textured vehicle mesh (diffuse / normal / specular / gloss maps), 640x480, camera {0,5,65}, back-face culling, point sampling",
                         "vehicle1080": "BASELINE config 2: the reference's bundled textured vehicle mesh (diffuse / normal / specular / gloss maps), 1920x1080, 16x16 tiles, back-face culling, point sampling"}[name]
                        + (f" ({scene.texture_source} texture maps)" if hasattr(scene, "texture_source") else ""),
            "triangles": scene.n_tris, "vertices": scene.n_vertices, "resolution": f"{scene.width}x{scene.height}",
            "l2_policy": "inputs larger than L2 (static position / triangle-reference / vertex streams 530 MB; ~0.9 GB of DRAM traffic per frame)" if name == "spheres10m" else "working set is L2-resident by nature of the workload (textures + 128 triangles); framebuffer write-back 66 MB per frame"}


# ------------------------------------------------------------------------------------ GPU arm
def algorithmic_bytes(scene, visible, pairs, frags, covered_px, direct=False):
    """SURVEY.md section 8d / BASELINE.md section 3 contract figures.  direct: the frame resolved its small triangles in the
    setup kernel (B2R_OPT_DIRECT_MAX_AREA), so the consumption of the (triangle, tile) pairs and the depth test of the
    fragments are work of the SETUP stage; the raster stage keeps the resolve + write-back.  The frame total is the same."""
    V, T, Tv, P, N = scene.n_vertices, scene.n_tris, visible, pairs, scene.bins_x * scene.bins_y
    px = scene.width * scene.height
    moved = (frags * 4 + P * 52) if direct else 0
    return {
        "vertex": V * 128,
        "setup": T * 60 + Tv * 12 + moved,
        "binning": Tv * 120 + P * 4 + N * 16,
        # fill/ROP: F*4 (depth read) + Fp*8 (depth+colour write) + W*H*8 (write-back) + list consumption P*(4+48)
        "raster_shade": frags * 4 + covered_px * 8 + px * 8 + P * 52 - moved,
    }


def refuse_emulation():
    """B2R_LIBRARY selects A/B builds of the product library; the CPU emulation build of tools/emu is test infrastructure and
    is never measured."""
    from custom_gpgpu_rasterizer_b200 import _lib
    if hasattr(_lib.lib, "emu_enqueue_kernel"):
        raise SystemExit(f"bench.py: {_lib.LIB_PATH} is the CPU emulation build (tools/emu) - there is no CPU path to measure")


def run_b200(args, scene):
    import torch
    import torch.distributed as dist
    from custom_gpgpu_rasterizer_b200 import _lib
    from custom_gpgpu_rasterizer_b200.renderer import Renderer
    refuse_emulation()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    r = Renderer(scene.width, scene.height, device=local)
    sharded = world > 1 and args.gather == "sharded"
    # frame ring (B2R_OPT_FRAME_RING): frame f runs on ring stream f mod D with its own buffer set, so the front ends of D
    # consecutive frames overlap; measured best at D = 3 on one GPU and D = 4 with the sharded front end.  B2R_FRAME_RING in
    # the environment overrides (0 = the three-stream pipeline).
    ring = int(os.environ["B2R_FRAME_RING"]) if "B2R_FRAME_RING" in os.environ else (args.ring if args.ring >= 0 else (4 if sharded else 3))
    if world == 1 or sharded:
        r.set_option(_lib.OPT_FRAME_RING, ring)
    # sharded front end: every rank transforms only the vertices of its own primitive slice, the raster kernel transforms the
    # vertices of its pixels' winners itself (measured at 8 ranks: 9 370 against 7 958 frames/s); needs the frame ring
    if sharded and ring and "B2R_SHARD_SLICED_VERTEX" not in os.environ:
        r.set_option(_lib.OPT_SHARD_SLICED_VERTEX, 1)
    if not sharded:
        r.set_option(_lib.OPT_TILE_OWNER_COUNT, world)
        r.set_option(_lib.OPT_TILE_OWNER_RANK, rank)
    t0 = time.perf_counter()
    r.load_scene(scene)
    load_s = time.perf_counter() - t0
    stream = torch.cuda.Stream(device=dev)
    r.set_stream(stream.cuda_stream)
    from custom_gpgpu_rasterizer_b200.renderer import host_camera
    cam = host_camera(scene.cam_pos, scene.fov_deg, scene.width, scene.height, scene.near, scene.far)

    gather = None
    if world > 1:
        from custom_gpgpu_rasterizer_b200.multi_gpu import ShardedSortFirst, SortFirstGather
        gather = ShardedSortFirst(r, rank, world) if sharded else SortFirstGather(r, scene, rank, world, stream, mode=args.gather)

    prepared = r.prepare_camera(cam)          # the camera marshalled once: the frame loop is one ctypes call per frame

    def frame():
        r.render_prepared(prepared, scene.cull, scene.sample, scene.vis)
        if gather is not None:
            gather.run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also grows the bin-list buffer if needed)
    for _ in range(max(args.warmup, 3)):
        frame()
    r.sync()
    barrier()

    # ---- timed region: device time of K frames, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # clocks / throttle reasons are sampled (nvidia-smi, every 100 ms) from here to the end of the per-stage timing below:
    # the device-timed region, the end-to-end region and the stage timers, plus ~1 s of the same frames in between so that
    # the median is taken over enough samples under load (K frames alone last a few milliseconds)
    clocks = ClockSampler(local).__enter__()
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            frame()
        e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    counters = r.counters()
    # untimed: the same frames for about a second (a fixed number on every rank, derived from the measured frame time)
    for _ in range(min(4000, int(1000.0 / max(total_ms / args.steps, 0.05)))):
        frame()
    r.sync()
    barrier()
    launches_per_frame = counters["kernels_launched"] + (gather.launches_per_frame if gather else 0)

    # ---- the same frames WITH the in-frame list sort (B2R_OPT_SORT_BINS = 1): the image does not depend on the list
    #      order (order-independent depth resolve) and b2r_read_bins canonicalises on demand, so the library default leaves
    #      the lists in scatter order; auxiliary number
    sorted_fps, sorter_counters = None, None
    if world == 1:
        r.set_option(_lib.OPT_SORT_BINS, 1)
        for _ in range(3):
            frame()
        r.sync()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            u0.record(stream)
            for _ in range(args.steps):
                frame()
            u1.record(stream)
        torch.cuda.synchronize()
        sorted_fps = args.steps / (u0.elapsed_time(u1) / 1e3)
        sorter_counters = r.counters()
        r.set_option(_lib.OPT_SORT_BINS, 0)
        frame()
        r.sync()

    # ---- host cost of enqueueing one frame (ring of 4 frames in flight: the first three calls after a sync never block)
    r.sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        frame()
    host_enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    r.sync()

    # ---- end to end through the public API: per-frame constants H2D + render + framebuffer D2H to host memory
    import zlib
    frame_bytes = scene.width * scene.height * 4
    crc_host = crc_dev = None
    distributed = sharded and args.present != "central"
    if distributed:
        # Distributed Present (N > 1): the host frame lives in memory all ranks share (a /dev/shm mapping, two frames) and every
        # rank delivers its part over its own PCIe link.  "band" (default): the frame is composed on rank 0 over NVLink as in
        # the device-timed region and rank r copies the r-th band of rows out of rank 0's memory with its own copy engine
        # (b2r_present_band_async).  "tiles": every rank keeps the colours of its tiles and stores them into the host frame
        # from a kernel (b2r_present_tiles_async); nothing crosses NVLink.
        tiles_mode = args.present == "tiles"
        shm = SharedHostFrames(f"b2r_bench_{os.environ.get('MASTER_PORT', '0')}", 2 * frame_bytes, rank, barrier)
        host_fb = shm.array.reshape(2, scene.height, scene.width)
        r.register_host_frame(shm.array)
        band_rows = None
        if tiles_mode:
            present = r.present_tiles_async
        else:
            # the GPUs of a box do not always reach the host equally fast (measured on this pool's 8-GPU boxes: four links at
            # ~16 GB/s, four at ~48 GB/s when all copy at once), so the bands are sized by a short calibration: every rank
            # copies 1/N of a frame into the shared frame 30 times, all ranks at once, and gets rows in proportion to its rate
            band_rows = calibrate_bands(torch, dist, dev, shm.array, scene, rank, world, barrier)
            present = lambda off: r.present_rows_async(off, band_rows[rank][0], band_rows[rank][1])      # noqa: E731
        if tiles_mode:
            r.set_option(_lib.OPT_SHARD_LOCAL_COLOUR, 1)
        for i in range(2):
            frame()
            present((i & 1) * frame_bytes)
        r.present_wait()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            frame()
            present((i & 1) * frame_bytes)
        r.present_wait()
        barrier()
        e2e_s = time.perf_counter() - t0
        if rank == 0:
            crc_host = zlib.crc32(host_fb[(args.steps - 1) & 1].tobytes())
        barrier()
        if tiles_mode:
            r.set_option(_lib.OPT_SHARD_LOCAL_COLOUR, 0)
        r.register_host_frame(None)
        host_fb = None
        shm.close()
        frame()                                   # a frame composed on rank 0's device again, for the device-side checksum
    else:
        pinned = [torch.empty((scene.height, scene.width), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        host_fb = [p.numpy().view(np.uint32) for p in pinned]
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            frame()
            if rank == 0:
                # Present(): device->host copy of the (gathered) frame into pinned memory; pipelined - the copy of
                # frame i overlaps the rendering of frame i+1 (two colour surfaces, two host buffers)
                if world == 1 or sharded:
                    r.present_async(host_fb[i & 1])
                else:
                    r.framebuffer(out=host_fb[i & 1])
        if rank == 0 and (world == 1 or sharded):
            r.present_wait()
        else:
            r.sync()
        barrier()
        e2e_s = time.perf_counter() - t0
        if rank == 0:
            crc_host = zlib.crc32(host_fb[(args.steps - 1) & 1].tobytes())
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    # correctness token of this run: CRC-32 of the frame the end-to-end region delivered into host memory and of the frame on
    # the device (N > 1: the frame the ranks composed on rank 0); compared below with a frame rank 0 renders on its own
    if rank == 0:
        crc_dev = zlib.crc32(r.framebuffer().tobytes())
    else:
        r.sync()
    barrier()

    # ---- per-stage device times over K more frames (CUDA events recorded between the stages on the same stream)
    r.set_option(_lib.OPT_STAGE_TIMERS, 1)
    r.stage_ms_mean(reset=True)
    for _ in range(args.steps):
        frame()
    r.sync()
    stages, n_timed = r.stage_ms_mean(reset=True)
    r.set_option(_lib.OPT_STAGE_TIMERS, 0)
    clocks.__exit__()
    stages_per_rank = None
    if world > 1:
        stages_per_rank = [None] * world
        dist.all_gather_object(stages_per_rank, {k: round(v, 4) for k, v in stages.items()})

    # ---- workload counters for the roofline (instrumented frame, outside every timed region)
    r.set_option(_lib.OPT_COUNT_FRAGMENTS, 1)
    frame()
    r.sync()
    c2 = r.counters()
    r.set_option(_lib.OPT_COUNT_FRAGMENTS, 0)
    total_tris, visible = r.stats()
    pairs, frags = counters["pairs"], c2["fragments"]
    if world > 1:
        # every rank counted its own tiles (and, with the sharded front end, its own primitives)
        agg = torch.tensor([pairs, frags, visible if sharded else 0], device=dev, dtype=torch.int64)
        dist.all_reduce(agg)
        pairs, frags = int(agg[0].item()), int(agg[1].item())
        if sharded:
            visible = int(agg[2].item())
    barrier()

    # ---- view-parallel mode (N > 1): every rank renders complete, independent frames - no data-path collective
    view_parallel = None
    if world > 1:
        r.sync()
        gather.close()
        r.bind_framebuffer(None)
        r.set_option(_lib.OPT_PACKED_OUTPUT, 0)
        r.set_option(_lib.OPT_TILE_OWNER_COUNT, 1)
        r.set_option(_lib.OPT_TILE_OWNER_RANK, 0)
        for _ in range(3):
            r.render(cam, scene.cull, scene.sample, scene.vis, sync=False)
        r.sync()
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            v0.record(stream)
            for _ in range(args.steps):
                r.render(cam, scene.cull, scene.sample, scene.vis, sync=False)
            v1.record(stream)
        barrier()
        vms = torch.tensor([v0.elapsed_time(v1)], device=dev, dtype=torch.float64)
        dist.all_reduce(vms, op=dist.ReduceOp.MAX)
        view_parallel = {"value": world * args.steps / (float(vms.item()) / 1e3), "unit": "frames/s", "scaling": "weak",
                         "note": "independent views: every rank renders whole frames of the same scene, no communication; max over ranks"}
        gather = None
    depth = r.depth()
    covered_px = int((depth != 0x7FFFFFFF).sum())
    # the reference frame of the checksum: rank 0 alone, whole frame, in this same process (N > 1: rendered by the
    # view-parallel leg above, after the sharded group was released)
    single = None
    if rank == 0:
        if world == 1:
            r.render(cam, scene.cull, scene.sample, scene.vis, sync=True)
        s_total, s_visible = r.stats()
        single = {"framebuffer_crc32": zlib.crc32(r.framebuffer().tobytes()), "visible_triangles": s_visible,
                  "tile_pairs": r.counters()["pairs"], "covered_pixels": covered_px}
    barrier()

    if rank != 0:
        if gather is not None:
            gather.close()
        r.close()
        if world > 1:
            dist.destroy_process_group()
        return

    fps = args.steps / (total_ms / 1e3)
    peak, peak_src = load_peaks()
    direct_area = r.get_option(_lib.OPT_DIRECT_MAX_AREA)
    ab = algorithmic_bytes(scene, visible, pairs, frags, covered_px, direct=direct_area > 0 and args.workload != "fill64")
    stage_of = {"vertex": "vertex", "setup": "setup", "binning": None, "raster_shade": "raster_shade"}
    per_stage = {}
    for k, b in ab.items():
        t = stages["bin_scan"] + stages["bin_scatter"] + stages["bin_sort"] if k == "binning" else stages[stage_of[k]]
        # N ranks: the bytes are the whole job's, the stage time is a rank's (the ranks work side by side), so the
        # denominator is the bandwidth of the N GPUs together
        per_stage[k] = {"ms": t, "algorithmic_bytes": b, "gbs": (b / 1e9) / (t / 1e3) if t > 0 else None,
                        "frac_of_hbm_peak": ((b / 1e9) / (t / 1e3)) / (peak * world) if t > 0 else None}
    dominant = max(per_stage, key=lambda k: per_stage[k]["ms"])
    roof = per_stage[dominant]
    # physical side of every stage, from the committed ncu capture of this build (profiles/ncu_traffic.json, written by
    # tools/ncu_traffic.py): DRAM bytes per launch, and the issue-slot picture of the stage's dominant kernel.  Physical
    # GB/s = those bytes / the stage time measured in THIS run.
    traffic, dominant_kernel = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            prof = json.load(f).get(args.workload, {})
        for k, st in per_stage.items():
            p = prof.get(k)
            if p is None or world != 1:
                continue
            nbytes = p["dram_bytes"] if isinstance(p, dict) else float(p)
            st["physical_dram_bytes"] = nbytes
            st["physical_gbs"] = (nbytes / 1e9) / (st["ms"] / 1e3) if st["ms"] > 0 else None
            st["physical_frac_of_hbm_peak"] = st["physical_gbs"] / peak if st["ms"] > 0 else None
            if isinstance(p, dict) and p.get("kernels"):
                kn, kd = max(p["kernels"].items(), key=lambda kv: kv[1]["time_us"])
                st["dominant_kernel"] = {"name": kn, "sm_throughput_pct": kd.get("sm_pct"), "dram_throughput_pct": kd.get("dram_pct"),
                                         "threads_per_instruction": kd.get("threads_per_inst"), "issue_active_pct": kd.get("issue_active_pct")}
        if world == 1 and dominant in prof:
            traffic = per_stage[dominant].get("physical_dram_bytes")
            dominant_kernel = per_stage[dominant].get("dominant_kernel")
    except Exception:      # noqa: BLE001
        pass
    phys_frac = per_stage[dominant].get("physical_frac_of_hbm_peak")
    # honest label: a stage that keeps DRAM below 40 % of the measured peak is not bandwidth-bound, whatever its effective figure
    bound = "hbm" if (phys_frac is None or phys_frac >= 0.4) else "issue/latency (DRAM < 40 % of peak: warp instruction issue and memory latency, see dominant_kernel)"

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_of(scene),
        "config": workload_config(args.workload, scene),
        "parallelism": ("single GPU" if world == 1 else f"sort-first, {world} ranks, tile owner = tile index mod {world}, " + {"nccl": "every rank bins the full mesh, NCCL gather of packed tiles to rank 0", "p2p": "every rank bins the full mesh, peer-to-peer ROP write-back into rank 0 over NVLink (CUDA IPC)", "sharded": "sharded front end: rank r transforms, sets up and bins primitive slice r; small triangles are resolved by that rank into its own key buffer and the touched 16x16 key blocks go to the tile owners over NVLink (direct path), the other triangles as list entries; owners merge, shade and store their colours straight into rank 0's surface; two flag barriers per frame, no collective"}[args.gather]),
        "mtris_per_s": scene.n_tris * fps / 1e6,
        "visible_triangles": visible, "tile_pairs": pairs, "fragments": frags, "covered_pixels": covered_px,
        "checksum": {"framebuffer_crc32": crc_dev, "host_framebuffer_crc32": crc_host, "single_gpu": single,
                     "match": bool(crc_dev == crc_host == single["framebuffer_crc32"] and visible == single["visible_triangles"] and pairs == single["tile_pairs"]),
                     "note": "CRC-32 of the 4K frame: on the device after the end-to-end region (N > 1: composed by all ranks on rank 0), as delivered into pinned host "
                             "memory by the end-to-end region, and of a whole frame rendered by rank 0 alone in this process; the visible-triangle and (triangle, tile) pair "
                             "counts summed over the ranks must equal the single-GPU frame's too.  A mismatch makes bench.py exit with status 3."},
        "stages_ms": stages, "stages": per_stage,
        "roofline": {"kernel": dominant, "bound": bound, "achieved": roof["gbs"], "peak": peak * world, "unit": "GB/s", "frac": roof["frac_of_hbm_peak"],
                     "traffic": traffic if world == 1 else None, "physical_frac": phys_frac, "dominant_kernel": dominant_kernel,
                     "peak_source": peak_src + ("" if world == 1 else f" x {world} GPUs"),
                     "note": "achieved = algorithmic bytes of SURVEY section 8d (the reference's record sizes: EFFECTIVE bandwidth, a stage that moves fewer bytes than "
                             "the reference's layout can exceed 1; physical DRAM bytes of the stage per launch are in `traffic`, from the committed ncu capture) / stage "
                             "time from CUDA events between the stages on the launching stream, averaged over %d frames" % n_timed},
        "e2e": {"value": args.steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": 64,
                "d2h_bytes_per_step": scene.width * scene.height * 4,
                "present": (args.present if distributed else "rank 0") if world > 1 else "single GPU",
                "band_rows_per_rank": [b[1] - b[0] for b in band_rows] if (distributed and band_rows) else None,
                "note": ("per frame: view-projection matrix from the host (zero-copy read; per-mesh constants are built on the device) + b2r_render + " +
                         ("DISTRIBUTED Present into one host frame all ranks share (/dev/shm mapping registered with cudaHostRegister): " +
                          ("every rank stores the tiles it owns from a kernel over its own PCIe link (b2r_present_tiles_async)" if args.present == "tiles" else
                           "rank r copies the r-th band of rows of the frame composed on rank 0 with its own copy engine over its own PCIe link (b2r_present_band_async)") +
                          ", d2h bytes = the whole frame summed over the ranks; " if distributed else
                          "device->host copy of the framebuffer into pinned memory (b2r_present_async, pipelined over two surfaces; RenderAuto = Render + Present), ") +
                         "wall clock over all frames incl. the final wait and a barrier across the ranks")},
        "frame_ring": ring if (world == 1 or sharded) else 0,
        "direct_max_area": direct_area,
        "gpu_launches": launches_per_frame * args.steps,
        "gpu_launches_per_step": launches_per_frame,
        "host_enqueue_ms_per_frame": host_enqueue_ms,
        "load_scene_s": load_s,
        "clocks": clocks.summary(),
    }
    if sorted_fps is not None:
        line["with_in_frame_list_sort"] = {"value": sorted_fps, "unit": "frames/s",
                                           "tiles_sent_to_the_cta_sorter_by_the_warp_sorter": sorter_counters["sort_fallback_tiles"], "tiles_beyond_the_sorter": sorter_counters["unsorted_tiles"],
                                           "note": "B2R_OPT_SORT_BINS=1: same framebuffer bit for bit, the tile lists are also left in ascending primitive-id order in memory; the default (0) leaves them in scatter order and canonicalises them when they are read (b2r_read_bins)"}
    if view_parallel is not None:
        line["view_parallel"] = view_parallel
        line["stages_ms_per_rank"] = stages_per_rank
    # auxiliary legs (rank 0 only, after every collective): a failure here is reported in the line, it does not cost the headline
    if world == 1 and not args.no_cpu_baseline:
        try:
            mean_s, _, kind = cpu_reference_arm(scene, 3, 1)
            line["cpu_baseline"] = {"value": 1.0 / mean_s, "unit": "frames/s", "cores": 1, "kind": kind,
                                    "sample": f"the full frame of this workload, 3 frames after 1 warm-up, {CPU_KIND_NOTE[kind]}, single-threaded like the reference",
                                    "mtris_per_s": scene.n_tris / mean_s / 1e6, "host_cores_available": os.cpu_count()}
        except Exception as e:      # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"FAILED: {type(e).__name__}: {e}"}
    if world == 1 and args.workload == "spheres10m" and not args.no_fill:
        try:
            line["fill_test"] = fill_test(r.__class__, _lib, torch, dev, args, peak)
        except Exception as e:      # noqa: BLE001
            line["fill_test"] = {"error": f"{type(e).__name__}: {e}"}
    r.close()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not line["checksum"]["match"]:
        sys.stderr.write(f"bench.py: CHECKSUM MISMATCH {line['checksum']}\n")
        sys.exit(3)


def fill_test(Renderer, _lib, torch, dev, args, peak):
    """BASELINE config 4 on the same GPU: fill/ROP-equivalent bandwidth (north_star: >= 50 % of the HBM roofline)."""
    from custom_gpgpu_rasterizer_b200.renderer import host_camera
    sc = make_scene("fill64")
    r = Renderer(sc.width, sc.height, device=dev.index)
    r.load_scene(sc)
    stream = torch.cuda.Stream(device=dev)
    r.set_stream(stream.cuda_stream)
    cam = host_camera(sc.cam_pos, sc.fov_deg, sc.width, sc.height, sc.near, sc.far)
    for _ in range(3):
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
    r.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    r.set_option(_lib.OPT_STAGE_TIMERS, 1)
    r.stage_ms_mean(reset=True)
    for _ in range(args.steps):
        r.render(cam, sc.cull, sc.sample, sc.vis, sync=False)
    r.sync()
    stages, _ = r.stage_ms_mean(reset=True)
    r.set_option(_lib.OPT_STAGE_TIMERS, 0)
    r.set_option(_lib.OPT_COUNT_FRAGMENTS, 1)
    r.render(cam, sc.cull, sc.sample, sc.vis, sync=True)
    frags = r.counters()["fragments"]
    r.close()
    px = sc.width * sc.height
    rop_bytes = frags * 4 + frags * 8 + px * 8          # back-to-front: every fragment passes depth (F = Fp)
    t = stages["raster_shade"]
    return {"workload": "BASELINE config 4: 64 stacked full-screen textured quads, 3840x2160, linear sampling", "texture": getattr(sc, "fill_texture", None),
            "frames_per_s": 1e3 / ms, "ms_per_frame": ms, "fragments": frags, "gfragments_per_s": frags / (t / 1e3) / 1e9,
            "raster_shade_ms": t, "rop_equivalent_bytes": rop_bytes, "rop_equivalent_gbs": (rop_bytes / 1e9) / (t / 1e3),
            "frac_of_hbm_peak": (rop_bytes / 1e9) / (t / 1e3) / peak, "physical_write_back_bytes": px * 8, "stages_ms": stages}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="spheres10m", choices=["spheres10m", "fill64", "spheres_small", "views172", "vehicle480", "vehicle1080"])
    ap.add_argument("--gather", default="sharded", choices=["nccl", "p2p", "sharded"], help="sort-first composition: NCCL gather of packed tiles, peer-to-peer ROP write-back over NVLink, or the sharded front end (b2r_shard_*)")
    ap.add_argument("--views-mode", default="parallel", choices=["parallel", "sortfirst"], help="--workload views172 at N > 1: independent views per rank (default) or every view rendered sort-first by all ranks")
    ap.add_argument("--present", default="band", choices=["band", "tiles", "central"], help="N > 1, end-to-end leg: distributed Present by row bands out of rank 0's surface (default) or by owned tiles, or through rank 0's PCIe link alone")
    ap.add_argument("--ring", type=int, default=-1, help="frame ring depth (B2R_OPT_FRAME_RING): 0 = three-stream pipeline, 2..4; default 3 on one GPU, 4 with the sharded front end")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fill", action="store_true")
    ap.add_argument("--ref-cuda", type=int, default=2, metavar="FRAMES", help="--impl reference only: also time the reference's own CUDA renderer (oracle/_ref/ref_render) on GPU 0 for this many frames (0 = skip)")
    ap.add_argument("--ref-cuda-timeout", type=int, default=240)
    args = ap.parse_args()
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "views172":
        if args.impl == "reference":
            raise SystemExit("the reference arm is defined on the default workload")
        return run_views(args)
    scene = make_scene(args.workload)
    if args.impl == "reference":
        run_reference(args, scene)
    else:
        run_b200(args, scene)


if __name__ == "__main__":
    main()
