#!/usr/bin/env python
"""Shift2d emulating a stride-2 depth-wise convolution (modules/shifts.py:85-89) on the cfg3 tensor: the fused operators
(shift + crop + 2x2 average pooling in one kernel forward; pooling adjoint applied while the gradient is staged backward)
against the two-step path (shift op + ATen avg_pool2d and their backwards).  CUDA events, graph replays."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401

dev = torch.device("cuda:0")
shape = (256, 256, 56, 56)
torch.manual_seed(0)
x = torch.randn(shape, device=dev)
gp = torch.randn(shape[0], shape[1], 28, 28, device=dev)
std = torch.tensor([0, 56, 0, 56, 0, 1], dtype=torch.int32)
ops = torch.ops.torchshifts


def graph_time(fn, reps=20):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr), torch.no_grad():
        fn()
    for _ in range(3):
        gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps):
        gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


for pad, active in ((0, False), (0, True), (3, True)):
    w = (torch.rand(shape[1], 2, device=dev) * 2 - 1) * 2
    f_fused = graph_time(lambda: ops._shift2d_avgpool2_forward(x, w, std, list(shape), pad, active))
    f_two = graph_time(lambda: torch.nn.functional.avg_pool2d(ops._shift2d_forward(x, w, std, list(shape), pad, active), 2, 2, ceil_mode=True))
    b_fused = graph_time(lambda: ops._shift2d_avgpool2_backward(gp, w, x, std, list(shape), pad, active))
    like = torch.empty(shape, device=dev)
    b_two = graph_time(lambda: ops._shift2d_backward(torch.ops.aten.avg_pool2d_backward(gp, like, [2, 2], [2, 2], [0, 0], True, True, None), w, x, std, pad, active))
    print(f"cfg3 stride-2 layer pad={pad} active={active}: forward fused {f_fused:7.1f} us | two-step {f_two:7.1f} us    backward fused {b_fused:7.1f} us | two-step {b_two:7.1f} us", flush=True)
