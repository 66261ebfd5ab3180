#!/usr/bin/env python
"""Where does a SMALL step go?  cfg3 at 32 images per GPU (the 8-GPU strong-scaling shard) on one GPU: CUDA-graph
replays of the forward only, the backward only (+ pass 2), both, and both with the in-kernel exchange (one-rank
peer group), each as us per replay (CUDA events over a loop of replays, after warm-up).
    python tools/step_profile.py [N] [--eager]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.functional import shift2d_func  # noqa: E402
from torchshifts.host import GraphedShiftStep  # noqa: E402
from torchshifts.sharded import FusedGradWeightAllReduce  # noqa: E402

N = next((int(a) for a in sys.argv[1:] if a.isdigit()), 32)
dev = torch.device("cuda:0")
shape = (N, 256, 56, 56)
torch.manual_seed(0)
x = torch.randn(shape, device=dev); g = torch.randn(shape, device=dev)
w = torch.rand(256, 2, device=dev) * 2 - 1


def timed(body, reps=200):
    for _ in range(10):
        body()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        body()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


step = GraphedShiftStep(x, w, g, 0, False, split=True)
one = GraphedShiftStep(x, w, g, 0, False, split=False)
fused = FusedGradWeightAllReduce(capacity=4096, device=dev)
with fused:
    stepf = GraphedShiftStep(x, w, g, 0, False, split=False)
elems = x.numel()
floor = elems * 20 / 6541.8e9 * 1e6
print(f"cfg3 shard of {N} images: HBM floor at the measured peak {floor:.1f} us (fwd {elems * 8 / 6541.8e9 * 1e6:.1f}, bwd {elems * 12 / 6541.8e9 * 1e6:.1f})")
print(f"forward graph only          {timed(step.replay_forward):7.1f} us")
print(f"backward graph only         {timed(step.replay_backward):7.1f} us")
print(f"forward + backward, 2 graphs {timed(step.replay):7.1f} us")
print(f"forward + backward, 1 graph  {timed(one.replay):7.1f} us")
print(f"1 graph + in-kernel exchange (1 rank) {timed(stepf.replay):7.1f} us")
if "--eager" in sys.argv:
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)

    def eager():
        xr.grad = None; wr.grad = None
        shift2d_func(xr, wr, 0, False).backward(g)
    print(f"eager public API            {timed(eager, 50):7.1f} us")
