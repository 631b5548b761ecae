"""bench.py without a GPU: the reference arm (the CPU port, timed on the host cores) prints the contract's JSON line, the
algorithmic-byte model is SURVEY section 8d's, and the GPU arm fails loudly instead of falling back."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("b2r_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--workload", "spheres_small", "--steps", "2", "--warmup", "1", "--ref-cuda", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    bench = _bench_module()
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-6
    # the reference's own CPU renderer when oracle/build_ref_erenderer.sh has built it (this container, and the GPU box, to
    # which oracle/_ref travels), else the restatement pinned against it
    built = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libere_ref.so"))
    assert d["cpu_baseline"]["kind"] == ("reference" if built else "port")
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["triangles"] == 400_000 and d["config"]["resolution"] == "3840x2160"


def test_reference_arm_can_be_forced_onto_the_restatement():
    res = _run("--impl", "reference", "--workload", "spheres_small", "--steps", "1", "--warmup", "0", "--ref-cuda", "0", env={"B2R_CPU_BASELINE_PORT": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    assert json.loads(res.stdout)["cpu_baseline"]["kind"] == "port"


def test_reference_arm_on_baseline_config_1():
    """BASELINE configs[0]: the reference's CPU software path on its bundled textured mesh at 640x480."""
    res = _run("--impl", "reference", "--workload", "vehicle480", "--steps", "2", "--warmup", "1", "--ref-cuda", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout)
    assert d["config"]["triangles"] == 11638 and d["config"]["resolution"] == "640x480" and "bundled vehicle mesh" in d["data"]
    assert 1.0 < d["ms_per_step"] < 2000.0


def test_reference_arm_other_ranks_exit_without_work():
    res = _run("--impl", "reference", "--workload", "spheres_small", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_algorithmic_bytes_are_surveys_figures():
    """SURVEY.md section 8d at config 3: vertex 663 MB, setup ~660 MB, binning ~630 MB; config 4 fill 6.436 GB."""
    bench = _bench_module()

    class S:
        n_vertices, n_tris, bins_x, bins_y, width, height = 5_176_250, 10_000_000, 240, 135, 3840, 2160
    ab = bench.algorithmic_bytes(S, visible=4_909_458, pairs=7_117_055, frags=3_089_508, covered_px=3_089_508)
    assert ab["vertex"] == 5_176_250 * 128 == 662_560_000
    assert ab["setup"] == 10_000_000 * 60 + 4_909_458 * 12
    assert ab["binning"] == 4_909_458 * 120 + 7_117_055 * 4 + 32_400 * 16
    assert ab["raster_shade"] == 3_089_508 * 12 + 3840 * 2160 * 8 + 7_117_055 * 52
    f = 64 * 3840 * 2160
    assert f * 12 + 3840 * 2160 * 8 == 6_436_454_400          # the contract figure of the fill test (F = Fp = 64 W H)


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_gpu_arm_fails_loudly_without_a_gpu():
    res = _run("--workload", "spheres_small", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-fill")
    assert res.returncode != 0 and res.stdout.strip() == "", "no CPU fallback: the GPU arm must not print a line without a device"


# ---------------------------------------------------------------------------------------------------------------------
# Dry run of the GPU arm's Python plumbing with a fake device: everything bench.py does around the renderer (timing
# brackets, per-stage roofline, end-to-end leg, auxiliary legs, multi-rank reductions, the JSON line) executes on CPU with
# the ctypes Renderer and torch.cuda replaced by stand-ins.  It cannot measure anything; it catches a typo that would
# otherwise cost the bench line at the round-end gate.
class _FakeRenderer:
    instances = []

    def __init__(self, width, height, device=0):
        self.w, self.h, self.opts, self.frames = width, height, {}, 0
        _FakeRenderer.instances.append(self)

    def set_option(self, o, v):
        self.opts[o] = v

    def get_option(self, o):
        return self.opts.get(o, 0)

    def load_scene(self, scene):
        self.scene = scene

    def set_stream(self, ptr):
        pass

    def render(self, cam, cull=0, sample=0, vis=0, clear_rgb=None, sync=True):
        assert cam["view_proj"].shape == (16,)
        self.frames += 1

    def prepare_camera(self, cam):
        return cam

    def render_prepared(self, prep, cull=0, sample=0, vis=0):
        self.render(prep, cull, sample, vis, sync=False)

    def sync(self):
        pass

    def counters(self):
        return {"pairs": 7_117_055, "kernels_launched": 10, "tiles_owned": 32400, "num_bins": 32400, "unsorted_tiles": 0, "big_triangles": 0,
                "fragments": 3_089_508, "sort_fallback_tiles": 0}

    def stats(self):
        return self.scene.n_tris, self.scene.n_tris // 2

    def stage_ms_mean(self, reset=True):
        from custom_gpgpu_rasterizer_b200 import _lib
        ms = dict(zip(_lib.STAGE_NAMES, (0.04, 0.10, 0.014, 0.055, 0.066, 0.238, 0.513, 0.003, 0.016)))
        return ms, 20

    def present_async(self, out):
        assert out.shape == (self.h, self.w) and out.dtype == np.uint32
        out[:] = self.framebuffer()

    def present_wait(self):
        pass

    def register_host_frame(self, frame):
        self.host_frame = frame

    def present_band_async(self, byte_offset=0):
        self.present_tiles_async(byte_offset)

    def present_rows_async(self, byte_offset, row_begin, row_end):
        assert 0 <= row_begin <= row_end <= self.h
        self.present_tiles_async(byte_offset)

    def present_tiles_async(self, byte_offset=0):
        n = self.w * self.h
        self.host_frame[byte_offset // 4: byte_offset // 4 + n] = self.framebuffer().reshape(-1)

    def framebuffer(self, out=None):
        return np.zeros((self.h, self.w), np.uint32)

    def depth(self):
        d = np.full((self.h, self.w), 0x7FFFFFFF, np.int32)
        d[: self.h // 2] = 5
        return d

    def bind_framebuffer(self, ptr):
        pass

    def close(self):
        pass


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 9.4          # ms for the K frames


class _FakeShard:
    launches_per_frame = 0

    def __init__(self, renderer, rank, world, *a, **k):
        pass

    def run(self):
        pass

    def close(self):
        pass


def _fake_device(monkeypatch, world=1):
    import contextlib
    import types

    import numpy  # noqa: F401
    import torch
    import torch.distributed as dist
    from custom_gpgpu_rasterizer_b200 import multi_gpu, renderer
    cpu = torch.device("cpu")
    monkeypatch.setattr(renderer, "Renderer", _FakeRenderer)
    monkeypatch.setattr(multi_gpu, "ShardedSortFirst", _FakeShard)
    monkeypatch.setattr(multi_gpu, "SortFirstGather", _FakeShard)
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    monkeypatch.setattr(torch.cuda, "set_device", lambda i: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(dist, "all_gather_object", lambda out, obj: out.__setitem__(slice(None), [obj] * len(out)))
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")


CONTRACT_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                 "config", "roofline", "e2e", "gpu_launches", "clocks"}


def _dry_run(monkeypatch, capsys, argv, world=1):
    _fake_device(monkeypatch, world)
    bench = _bench_module()
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_gpu_arm_dry_run_single_rank(monkeypatch, capsys):
    d = _dry_run(monkeypatch, capsys, ["--steps", "20", "--warmup", "5"])
    assert CONTRACT_KEYS <= set(d) and "cpu_baseline" in d and "fill_test" in d
    assert d["n_gpus"] == 1 and d["config"]["triangles"] == 10_000_000 and "config 3" in d["config"]["workload"]
    assert abs(d["value"] - 20 / 9.4e-3) < 1e-6 and d["gpu_launches"] == 10 * 20
    assert d["checksum"]["match"] is True and d["checksum"]["framebuffer_crc32"] == d["checksum"]["host_framebuffer_crc32"]
    r = d["roofline"]
    assert r["kernel"] == "raster_shade" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"]
    # the label follows the PHYSICAL fraction (ncu DRAM bytes / this run's stage time): below 40 % of the HBM peak it is not "hbm"
    assert (r["bound"] == "hbm") == (r["physical_frac"] >= 0.4) and abs(r["physical_frac"] - r["traffic"] / 1e9 / (d["stages"]["raster_shade"]["ms"] / 1e3) / r["peak"]) < 1e-9
    assert d["e2e"]["d2h_bytes_per_step"] == 3840 * 2160 * 4 and d["e2e"]["value"] > 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0
    assert "error" not in d["fill_test"] and d["fill_test"]["rop_equivalent_bytes"] == 3_089_508 * 12 + 3840 * 2160 * 8


def test_gpu_arm_dry_run_eight_ranks(monkeypatch, capsys):
    d = _dry_run(monkeypatch, capsys, ["--gpus", "8", "--steps", "10", "--warmup", "3", "--no-cpu-baseline"], world=8)
    assert CONTRACT_KEYS <= set(d) and d["n_gpus"] == 8 and "view_parallel" in d and len(d["stages_ms_per_rank"]) == 8
    assert "sharded front end" in d["parallelism"] and "parallelism" not in d["config"]
    assert d["checksum"]["match"] is True
    r = d["roofline"]
    assert abs(r["peak"] / 8 - _bench_module().load_peaks()[0]) < 1e-6 and r["traffic"] is None      # N GPUs' bandwidth; ncu traffic is a 1-GPU capture


@pytest.mark.parametrize("mode,world", [("parallel", 1), ("parallel", 4), ("sortfirst", 4)])
def test_views_workload_dry_run(monkeypatch, capsys, mode, world):
    d = _dry_run(monkeypatch, capsys, ["--workload", "views172", "--gpus", str(world), "--steps", "8", "--views-mode", mode], world=world)
    assert d["unit"] == "views/s" and d["config"]["triangles"] == 172 * 11638
    assert abs(d["value"] - (1 if mode == "sortfirst" else world) * 8 / 9.4e-3) < 1e-6
    assert ("sort-first" in d["parallelism"]) == (mode == "sortfirst")


def test_vehicle_workload_dry_run(monkeypatch, capsys):
    d = _dry_run(monkeypatch, capsys, ["--workload", "vehicle1080", "--steps", "5", "--no-cpu-baseline"])
    assert d["config"]["triangles"] == 11638 and d["config"]["resolution"] == "1920x1080" and "bundled vehicle mesh" in d["data"]
