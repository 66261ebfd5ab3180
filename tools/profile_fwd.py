#!/usr/bin/env python
"""cfg3 sparse and active forward back to back (for ncu comparisons)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
dev = torch.device("cuda:0")
torch.manual_seed(0)
shape = (256, 256, 56, 56)
x = torch.randn(shape, device=dev)
w = torch.rand(256, 2, device=dev) * 2 - 1
borders = torch.tensor([0, 56, 0, 56, 0, 1], dtype=torch.int32)
fwd = torch.ops.torchshifts._shift2d_forward
with torch.no_grad():
    for _ in range(3):
        y = fwd(x, w, borders, list(shape), 0, False)
        y = fwd(x, w, borders, list(shape), 0, True)
torch.cuda.synchronize()
