#!/usr/bin/env python
"""One forward + backward of a named workload through the registered ops, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:k_staged -s 4 -c 2 -o gpurun_out/prof python tools/profile_step.py cfg3
(two warm-up steps, then one profiled step; prints nothing but the kernel path)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.extension import native  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape, dim, pad, active = {"cfg3": ((256, 256, 56, 56), 2, 0, False), "cfg2": ((64, 512, 4096), 1, 2, True),
                           "cfg4": ((32, 128, 16, 56, 56), 3, 0, True), "cfg1": ((8, 64, 32, 32), 2, 0, False),
                           "cfg4r": ((32, 128, 16, 56, 56), 3, 3, True), "cfg3r": ((256, 256, 56, 56), 2, 3, False),
                           "cfg2z": ((64, 512, 4096), 1, 0, True), "cfg3n32": ((32, 256, 56, 56), 2, 0, False),
                           "cfg3ra": ((256, 256, 56, 56), 2, 3, True), "cfg5": ((256, 256, 56, 56), 2, 0, False), "cfg5cl": ((256, 256, 56, 56), 2, 0, False), "cfg4": ((32, 128, 16, 56, 56), 3, 0, True), "cfg2h": ((64, 512, 4096), 1, 2, True), "cfg4b": ((32, 128, 16, 56, 56), 3, 1, True)}[cfg]
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(shape, device=dev)
g = torch.randn(shape, device=dev)
w = torch.rand(shape[1], dim, device=dev) * 2 - 1
sp = list(shape[2:]) + [1] * (3 - dim)
borders = torch.tensor([0, sp[0], 0, sp[1], 0, sp[2]], dtype=torch.int32)
fwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_forward")
bwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_backward")
with torch.no_grad():
    if cfg in ("cfg5", "cfg5cl"):
        from torchshifts.quantized.modules.shifts import quantize_shift_weights
        xq = torch.quantize_per_tensor(torch.rand(shape, device=dev), 1 / 255., -128, torch.qint8)
        qw = quantize_shift_weights(w * 3)
        if cfg == "cfg5cl":
            xq = xq.contiguous(memory_format=torch.channels_last)
        for _ in range(steps):
            y = fwd(xq, qw, borders, list(shape), pad, False)
    else:
        if cfg == "cfg2h":
            x, g, w = x.bfloat16(), g.bfloat16(), w.bfloat16()
        for _ in range(steps):
            y = fwd(x, w, borders, list(shape), pad, active)
            gi, gw = bwd(g, w, x, borders, pad, active)
torch.cuda.synchronize()
print("kernel path", native().lib.ts_last_kernel_path())
