#!/usr/bin/env python
"""Backward (and forward) time of cfg3 shards of N images as a function of N and of the images per work unit
(chunk_planes): separates the fixed per-launch cost from the per-unit cost.  One GPU.
    python tools/small_step_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.extension import native  # noqa: E402

lib = native().lib
dev = torch.device("cuda:0")
fwd = torch.ops.torchshifts._shift2d_forward
bwd = torch.ops.torchshifts._shift2d_backward
borders = torch.tensor([0, 56, 0, 56, 0, 1], dtype=torch.int32)


def graph_time(fn, reps=200):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        fn()
    for _ in range(10):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


for N in (8, 16, 32, 64, 128, 256):
    shape = (N, 256, 56, 56)
    torch.manual_seed(0)
    x = torch.randn(shape, device=dev); g = torch.randn(shape, device=dev)
    w = torch.rand(256, 2, device=dev) * 2 - 1
    line = f"N={N:4d}:"
    for cp in (0, 1, 2, 4, 8):
        if cp > N:
            continue
        assert lib.ts_set_tuning(f"chunk_planes={cp}".encode()) == 0
        tf = graph_time(lambda: fwd(x, w, borders, list(shape), 0, False))
        tb = graph_time(lambda: bwd(g, w, x, borders, 0, False))
        line += f"  cp={cp}: fwd {tf:6.1f} bwd {tb:6.1f} us"
    lib.ts_set_tuning(b"chunk_planes=0")
    n = x.numel()
    print(line + f"   [floor fwd {n * 8 / 6541.8e3:.1f} bwd {n * 12 / 6541.8e3:.1f}]", flush=True)
    del x, g
