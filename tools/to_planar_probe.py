"""cfg3-sized fp32 channels-last tensor through ts_nhwc_to_nchw (the float path's layout pass): timing / ncu target.
Usage: python tools/to_planar_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts._ops import _dense  # noqa: E402

x = torch.randn(256, 256, 56, 56, device="cuda").contiguous(memory_format=torch.channels_last)
for _ in range(3):
    y = _dense(x)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    y = _dense(x)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"ts_nhwc_to_nchw fp32 {tuple(x.shape)}: {ms:.3f} ms  {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s")
assert torch.equal(y, x.contiguous())
