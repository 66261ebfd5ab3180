#!/usr/bin/env python
"""Sweep the host-buffer pipeline's chunk size / ring depth on cfg3 (e2e metric of bench.py) and measure
the raw pinned-memory copy ceilings of the box for comparison."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.host import HostShift2dPipeline  # noqa: E402

dev = torch.device("cuda:0")
N, C, H, W = 256, 256, 56, 56
w = torch.rand(C, 2, device=dev) * 2 - 1
# raw copy ceilings
a = torch.empty(N, C, H, W, pin_memory=True); b = torch.empty(N, C, H, W, pin_memory=True)
da = torch.empty(N, C, H, W, device=dev); db = torch.empty(N, C, H, W, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for label, both in (("H2D only", False), ("H2D + D2H concurrently", True)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                b.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"{label}: {a.numel() * 4 / dt / 1e9:.1f} GB/s per direction", flush=True)
del a, b, da, db
for chunk, slots in ((16, 3), (8, 3), (32, 3), (64, 3), (16, 2), (16, 4), (32, 4), (4, 4)):
    pipe = HostShift2dPipeline(N, C, H, W, device=dev, chunk=chunk, slots=slots)
    pipe.x_host.normal_(); pipe.g_host.normal_()
    for _ in range(2):
        pipe.forward_backward(w, 0, False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        pipe.forward_backward(w, 0, False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunk={chunk:3d} slots={slots}: {dt * 1e3:6.2f} ms/step  {N * C * H * W * 20 / dt / 1e9:6.1f} GB/s (algorithmic), "
          f"{pipe.h2d_bytes / dt / 1e9:.1f} GB/s per PCIe direction", flush=True)
    del pipe
