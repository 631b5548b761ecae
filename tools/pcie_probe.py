#!/usr/bin/env python
"""Pinned-memory D2H / H2D bandwidth of the box (explains the e2e ceiling: 33 MB framebuffer per frame)."""
import json, time, torch
dev = torch.device("cuda:0")
out = {}
for mb in (33, 256):
    n = mb * 1024 * 1024
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        out[f"{name}_{mb}MB_GBs"] = round(n * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
