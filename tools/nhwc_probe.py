"""cfg5 in channels-last through the native NHWC kernel: a few launches for ncu / timing.
Usage: python tools/nhwc_probe.py [reps] [shift spread]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402
from torchshifts.quantized.functional import shift2d_quantized  # noqa: E402
from torchshifts.quantized.modules.shifts import quantize_shift_weights  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.quantize_per_tensor(torch.rand(256, 256, 56, 56, device=dev), 1 / 255., -128, torch.qint8)
xcl = x.contiguous(memory_format=torch.channels_last)
del x
spread = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
qw = quantize_shift_weights((torch.rand(256, 2, device=dev) * 2 - 1) * spread)
lib = torchshifts.extension.native().lib
rows_sweep = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
for variant, pad, rows in [(2, 0, r) for r in rows_sweep] + [(2, 3, 0), (1, 0, 0)]:
    assert lib.ts_set_tuning(b"nhwc_variant=%d,nhwc_ring_rows=%d" % (variant, rows)) == 0
    for _ in range(reps):
        y = shift2d_quantized(xcl, qw, pad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        y = shift2d_quantized(xcl, qw, pad)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"variant {variant} ring_rows {rows} pad {pad}: {ms:.3f} ms  {2 * xcl.numel() / ms / 1e6:.0f} GB/s  path {torchshifts.extension.native().lib.ts_last_kernel_path()}")
