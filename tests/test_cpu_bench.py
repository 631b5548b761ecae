"""bench.py without a GPU: the reference arm (the CPU port, timed on the host cores) prints the contract's JSON line, the
algorithmic-byte model is SURVEY section 8d's, and the GPU arm fails loudly instead of falling back."""
import importlib.util
import json
import os
import subprocess
import sys

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
