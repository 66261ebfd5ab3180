#!/usr/bin/env python
"""Sweep of the TMA-family knobs (ring depth, stage size, consumer warps) for cfg3 forward and backward, each timed as
CUDA-graph replays (CUDA events, 50 replays after warm-up).     python tools/tma_sweep.py [N]"""
import itertools
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.extension import native  # noqa: E402

lib = native().lib
dev = torch.device("cuda:0")
N = next((int(a) for a in sys.argv[1:] if a.isdigit()), 256)
active = "--active" in sys.argv
shape = (N, 256, 56, 56)
fwd = torch.ops.torchshifts._shift2d_forward
bwd = torch.ops.torchshifts._shift2d_backward
borders = torch.tensor([0, 56, 0, 56, 0, 1], dtype=torch.int32)
torch.manual_seed(0)
x = torch.randn(shape, device=dev); g = torch.randn(shape, device=dev)
w = torch.rand(256, 2, device=dev) * 2 - 1


def graph_time(fn, reps=50):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr), torch.no_grad():
        fn()
    for _ in range(5):
        gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


res = {"fwd": [], "bwd": []}
for stages, kb, warps in itertools.product((2, 3, 4, 5, 6), (0, 28, 42, 56), (0, 8, 10, 12, 14)):
    spec = f"tma_stages={stages},tma_stage_kb={kb},tma_warps={warps}"
    assert lib.ts_set_tuning(spec.encode()) == 0
    try:
        tf = graph_time(lambda: fwd(x, w, borders, list(shape), 0, active))
        tb = graph_time(lambda: bwd(g, w, x, borders, 0, active))
    except RuntimeError as e:
        print(spec, "failed", str(e)[:80]); continue
    res["fwd"].append((tf, spec)); res["bwd"].append((tb, spec))
for k in ("fwd", "bwd"):
    print(f"--- {k} (N={N}, active={active}): best 8")
    for t, spec in sorted(res[k])[:8]:
        print(f"  {t:7.1f} us  {spec}")
    print(f"  default-ish: " + "; ".join(f"{t:.1f} {s}" for t, s in res[k] if s.endswith("tma_stage_kb=0,tma_warps=0")))
