#!/usr/bin/env python
"""Time every BASELINE.json configuration (forward and backward separately, back-to-back launches,
CUDA events) and print algorithmic GB/s, the kernel family that served it and the fraction of the
measured HBM peak.   python tools/bench_configs.py [cfg1 cfg2 ...]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.extension import native  # noqa: E402
from torchshifts.quantized.modules.shifts import quantize_shift_weights  # noqa: E402

lib = native().lib
dev = torch.device("cuda:0")
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    PEAK = 6650.0
PADS = ["zeros", "border", "periodic", "reflect", "symmetric"]
CASES = {
    "cfg1": [("cfg1 2d sparse zeros f32", (8, 64, 32, 32), 0, False, torch.float32)],
    "cfg2": [("cfg2 1d active periodic f32", (64, 512, 4096), 2, True, torch.float32),
             ("cfg2 1d active periodic bf16", (64, 512, 4096), 2, True, torch.bfloat16)],
    "cfg3": [("cfg3 2d sparse zeros f32", (256, 256, 56, 56), 0, False, torch.float32),
             ("cfg3 2d active zeros f32", (256, 256, 56, 56), 0, True, torch.float32),
             ("cfg3 2d sparse reflect f32", (256, 256, 56, 56), 3, False, torch.float32),
             ("cfg3 2d active reflect f32", (256, 256, 56, 56), 3, True, torch.float32)],
    "cfg4": [(f"cfg4 3d active {PADS[p]} f32", (32, 128, 16, 56, 56), p, True, torch.float32) for p in range(5)] +
            [("cfg4 3d sparse reflect f32", (32, 128, 16, 56, 56), 3, False, torch.float32)],
    "cfg5": [("cfg5 2d qint8 zeros", (256, 256, 56, 56), 0, False, torch.qint8),
             ("cfg5 2d quint8 zeros", (256, 256, 56, 56), 0, False, torch.quint8)],
}


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


names = {0: "none", 1: "generic", 2: "staged", 3: "tma", 4: "nhwc", 5: "halo", 6: "flat"}
want = [a for a in sys.argv[1:] if a in CASES] or list(CASES)
for key in want:
    for label, shape, pad, active, dtype in CASES[key]:
        dim = len(shape) - 2
        torch.manual_seed(0)
        sp = list(shape[2:]) + [1] * (3 - dim)
        borders = torch.tensor([0, sp[0], 0, sp[1], 0, sp[2]], dtype=torch.int32)
        fwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_forward")
        bwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_backward")
        w = torch.rand(shape[1], dim, device=dev) * 2 - 1
        n = 1
        for s in shape:
            n *= s
        with torch.no_grad():
            if dtype in (torch.qint8, torch.quint8):
                x = torch.quantize_per_tensor(torch.rand(shape, device=dev), 1 / 255., -128 if dtype == torch.qint8 else 0, dtype)
                qw = quantize_shift_weights(w * 3)
                tf = timeit(lambda: fwd(x, qw, borders, list(shape), pad, False))
                pf = lib.ts_last_kernel_path()
                print(f"{label:34s} fwd {tf:7.3f} ms {n * 2 / tf / 1e6:6.0f} GB/s ({n * 2 / tf / 1e6 / PEAK:4.0%}) [{names[pf]}]", flush=True)
                # the same tensor in channels-last: native NHWC kernel vs the three-pass route it replaces
                xcl = x.contiguous(memory_format=torch.channels_last)
                tf = timeit(lambda: fwd(xcl, qw, borders, list(shape), pad, False))
                pf = lib.ts_last_kernel_path()
                print(f"{label + ' NHWC':34s} fwd {tf:7.3f} ms {n * 2 / tf / 1e6:6.0f} GB/s ({n * 2 / tf / 1e6 / PEAK:4.0%}) [{names[pf]}]", flush=True)
                tc = timeit(lambda: fwd(xcl.contiguous(), qw, borders, list(shape), pad, False).contiguous(memory_format=torch.channels_last))
                print(f"{label + ' NHWC via NCHW':34s} fwd {tc:7.3f} ms {n * 2 / tc / 1e6:6.0f} GB/s ({n * 2 / tc / 1e6 / PEAK:4.0%}) [3 passes]", flush=True)
                del xcl
                continue
            x = torch.randn(shape, device=dev).to(dtype)
            g = torch.randn(shape, device=dev).to(dtype)
            w = w.to(dtype)
            es = x.element_size()
            tf = timeit(lambda: fwd(x, w, borders, list(shape), pad, active))
            pf = lib.ts_last_kernel_path()
            tb = timeit(lambda: bwd(g, w, x, borders, pad, active))
            pb = lib.ts_last_kernel_path()
        print(f"{label:34s} fwd {tf:7.3f} ms {n * 2 * es / tf / 1e6:6.0f} GB/s ({n * 2 * es / tf / 1e6 / PEAK:4.0%}) [{names[pf]}] | "
              f"bwd {tb:7.3f} ms {n * 3 * es / tb / 1e6:6.0f} GB/s ({n * 3 * es / tb / 1e6 / PEAK:4.0%}) [{names[pb]}]", flush=True)
        del x, g
        torch.cuda.empty_cache()
