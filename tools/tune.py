#!/usr/bin/env python
"""Sweep the staged-path tuning knobs (ts_set_tuning) on one workload and print a table.
    python tools/tune.py [cfg3|cfg2|cfg5|cfg1] [--quick]
Times forward and backward separately with CUDA events (median of 7 after 3 warm-ups)."""
import itertools
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.extension import native  # noqa: E402

lib = native().lib
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg3"
quick = "--quick" in sys.argv
torch.manual_seed(0)
if cfg == "cfg3":
    shape, dim, pad, active = (256, 256, 56, 56), 2, 0, False
elif cfg == "cfg2":
    shape, dim, pad, active = (64, 512, 4096), 1, 2, True
elif cfg == "cfg4":
    shape, dim, pad, active = (32, 128, 16, 56, 56), 3, 0, True
elif cfg == "cfg4r":
    shape, dim, pad, active = (32, 128, 16, 56, 56), 3, 3, True
elif cfg == "cfg3r":
    shape, dim, pad, active = (256, 256, 56, 56), 2, 3, False
else:
    shape, dim, pad, active = (8, 64, 32, 32), 2, 0, False
x = torch.randn(shape, device=dev)
g = torch.randn(shape, device=dev)
w = torch.rand(shape[1], dim, device=dev) * 2 - 1
borders = torch.tensor([0, shape[2], 0, shape[3] if dim > 1 else 1, 0, shape[4] if dim > 2 else 1], dtype=torch.int32)
fwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_forward")
bwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_backward")
elems = x.numel()


def timeit(fn, reps=20):
    """mean of `reps` back-to-back launches (one event pair around all of them, as bench.py times a step)"""
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(label):
    with torch.no_grad():
        tf = timeit(lambda: fwd(x, w, borders, list(shape), pad, active))
        path = lib.ts_last_kernel_path()
        tb = timeit(lambda: bwd(g, w, x, borders, pad, active))
    print(f"{label:46s} path={path} fwd {tf:7.3f} ms {elems * 8 / tf / 1e6:7.0f} GB/s | bwd {tb:7.3f} ms {elems * 12 / tb / 1e6:7.0f} GB/s | "
          f"fwd+bwd {tf + tb:7.3f} ms {elems * 20 / (tf + tb) / 1e6:7.0f} GB/s", flush=True)


print(f"workload {cfg} {shape} pad={pad} active={active}")
lib.ts_set_kernel_path(1)
run("generic")
lib.ts_set_kernel_path(0)
lib.ts_set_tuning(b"use_tma=0")
run("default tuning")
grid = [(st, kb, wp, ct) for st, kb, wp, ct in itertools.product((2, 3, 4, 6), (13, 26, 40, 56, 100), (8, 12, 16), (1, 2))]
if quick:
    grid = [(4, 48, 15, 1), (3, 64, 15, 1), (3, 72, 15, 1), (2, 100, 15, 1), (6, 32, 15, 1), (4, 48, 8, 1), (4, 48, 11, 1), (3, 72, 8, 1), (3, 72, 11, 1), (2, 110, 11, 1), (2, 110, 15, 1)]
if "--tma" in sys.argv:
    grid = []
for st, kb, wp, ct in grid:
    spec = f"stages={st},stage_kb={kb},warps={wp},ctas_per_sm={ct}"
    if lib.ts_set_tuning(spec.encode()) != 0:
        continue
    try:
        run(spec)
    except RuntimeError as e:
        print(spec, "failed:", str(e)[:120])
if "--tma" not in sys.argv:
    sys.exit(0)
print("--- TMA family")
lib.ts_set_tuning(b"use_tma=1")
run("tma default")
for st, kb, wp in itertools.product((3, 4, 6), (0, 28, 42, 84), (7, 9, 11, 13, 16, 19, 25)):
    spec = f"tma_stages={st},tma_stage_kb={kb},tma_warps={wp}"
    if lib.ts_set_tuning(spec.encode()) == 0:
        try:
            run(spec)
        except RuntimeError as e:
            print(spec, "failed:", str(e)[:120])
lib.ts_set_tuning(b"tma_stages=0,tma_stage_kb=0,tma_warps=0")
