"""Comparison helpers shared by the parity tests (product vs oracle vs reference)."""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RENDER = os.path.join(ROOT, "oracle", "_ref", "ref_render")

# tolerances of BASELINE.json north_star
DEPTH_RTOL = 1e-5          # depth buffers within 1e-5 relative
COLOUR_LSB = 1             # 8-bit colour within +-1 LSB per channel
COLOUR_MAX_FRACTION = 1e-3 # at most 0.1 % of pixels may differ by more, and only on shared edges


def bits(a: np.ndarray) -> np.ndarray:
    """Bit patterns of float32 data for exact comparison.  NaNs are canonicalised: the value class is part
    of the contract, the payload is not (x86 produces 0xFFC00000, the GPU 0x7FFFFFFF for the same operation;
    the bundled vehicle mesh carries NaN tangents from the reference parser's degenerate-uv triangles)."""
    f = np.ascontiguousarray(a, np.float32)
    u = f.view(np.uint32).copy()
    u[np.isnan(f)] = 0x7FC00000
    return u


def channels(fb: np.ndarray) -> np.ndarray:
    return np.stack([(fb >> 16) & 255, (fb >> 8) & 255, fb & 255], -1).astype(np.int32)


def colour_report(a: np.ndarray, b: np.ndarray, allowed_mask: np.ndarray | None = None) -> dict:
    """Per-pixel channel differences between two 0x00RRGGBB images."""
    d = np.abs(channels(a) - channels(b)).max(-1)
    bad = d > COLOUR_LSB
    rep = {"exact_fraction": float((d == 0).mean()), "within_1lsb_fraction": float((~bad).mean()),
           "bad_pixels": int(bad.sum()), "max_diff": int(d.max()), "n_pixels": int(d.size)}
    if allowed_mask is not None:
        rep["bad_outside_allowed"] = int((bad & ~allowed_mask).sum())
    return rep


def assert_colour_parity(a, b, allowed_mask=None, what=""):
    rep = colour_report(a, b, allowed_mask)
    assert rep["bad_pixels"] <= COLOUR_MAX_FRACTION * rep["n_pixels"], f"{what}: too many pixels beyond +-1 LSB: {rep}"
    if allowed_mask is not None:
        assert rep["bad_outside_allowed"] == 0, f"{what}: pixels beyond +-1 LSB away from shared edges / depth ties: {rep}"
    return rep


def assert_depth_parity(z_a, z_b, covered, what=""):
    za, zb = z_a[covered].astype(np.float64), z_b[covered].astype(np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        rel = np.abs(za - zb) / np.maximum(np.abs(zb), 1e-30)
    # non-finite depths (degenerate triangles: 0/0 weights) must agree in kind; equal infinities are equal
    rel = np.where(np.isnan(za) & np.isnan(zb), 0.0, np.where(za == zb, 0.0, np.where(np.isnan(rel), np.inf, rel)))
    assert rel.size == 0 or rel.max() <= DEPTH_RTOL, f"{what}: depth differs by {rel.max():.3e} relative (limit {DEPTH_RTOL})"
    return float(rel.max()) if rel.size else 0.0


def dilate(mask: np.ndarray, r: int = 1) -> np.ndarray:
    out = mask.copy()
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            sh = np.roll(np.roll(mask, dy, 0), dx, 1)
            out |= sh
    return out


def have_ref() -> bool:
    return os.path.exists(REF_RENDER)


def run_ref(scene, mode: str = "stages", timeout: int = 600) -> dict:
    """Run the unmodified reference renderer (oracle/_ref/ref_render) on a Scene (GPU box only)."""
    from custom_gpgpu_rasterizer_b200.scenefile import read_ref_dump
    with tempfile.TemporaryDirectory() as td:
        sp, op = os.path.join(td, "scene.b2rs"), os.path.join(td, "out.b2ro")
        with open(sp, "wb") as f:
            f.write(scene.to_blob())
        res = subprocess.run([REF_RENDER, mode, sp, op], capture_output=True, text=True, timeout=timeout)
        if res.returncode != 0:
            raise RuntimeError(f"ref_render failed ({res.returncode}): {res.stderr[-2000:]}")
        return read_ref_dump(op, scene)


def triples_of(scene, mesh_index: int, prim_ids: np.ndarray) -> np.ndarray:
    """Index triples (the reference's TriangleIdx) of local primitive ids of one mesh."""
    idx = scene.meshes[mesh_index].indices.reshape(-1, 3)
    return idx[prim_ids]


def sort_rows(a: np.ndarray) -> np.ndarray:
    if a.size == 0:
        return a.reshape(0, a.shape[-1] if a.ndim > 1 else 1)
    order = np.lexsort(a.T[::-1])
    return a[order]
