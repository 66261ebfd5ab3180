#!/usr/bin/env python
"""Time one BASELINE configuration (forward and backward, CUDA-graph replays) under a list of tuning specs.
    python tools/knob_sweep.py cfg4r "halo_stages=3" "halo_stages=4" ...
configs: cfg3 cfg3a cfg3r cfg3ra cfg4 cfg4r cfg4rs cfg5 cfg2 cfg2h (see CASES)."""
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
CASES = {"cfg1": ((8, 64, 32, 32), 0, False, torch.float32), "cfg3": ((256, 256, 56, 56), 0, False, torch.float32), "cfg3a": ((256, 256, 56, 56), 0, True, torch.float32),
         "cfg3r": ((256, 256, 56, 56), 3, False, torch.float32), "cfg3ra": ((256, 256, 56, 56), 3, True, torch.float32),
         "cfg4": ((32, 128, 16, 56, 56), 0, True, torch.float32), "cfg4r": ((32, 128, 16, 56, 56), 3, True, torch.float32),
         "cfg4rs": ((32, 128, 16, 56, 56), 3, False, torch.float32), "cfg4b": ((32, 128, 16, 56, 56), 1, True, torch.float32), "cfg5": ((256, 256, 56, 56), 0, False, torch.qint8), "cfg5cl": ((256, 256, 56, 56), 0, False, torch.qint8),
         "cfg2": ((64, 512, 4096), 2, True, torch.float32), "cfg2h": ((64, 512, 4096), 2, True, torch.bfloat16),
         "cfg2h8k": ((32, 512, 8192), 2, True, torch.bfloat16), "cfg2h2k": ((128, 512, 2048), 2, True, torch.bfloat16),
         "cfg3d": ((128, 256, 56, 56), 0, False, torch.float64), "cfg3da": ((128, 256, 56, 56), 3, True, torch.float64),
         "cfg4d": ((16, 128, 16, 56, 56), 3, True, torch.float64),
         "cfg2f2k": ((128, 512, 2048), 2, True, torch.float32), "cfg2f8k": ((32, 512, 8192), 2, True, torch.float32)}
shape, pad, active, dtype = CASES[sys.argv[1]]
specs = sys.argv[2:] or [""]
dim = len(shape) - 2
sp = list(shape[2:]) + [1] * (3 - dim)
borders = torch.tensor([0, sp[0], 0, sp[1], 0, sp[2]], dtype=torch.int32)
fwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_forward")
bwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_backward")
torch.manual_seed(0)
w = torch.rand(shape[1], dim, device=dev) * 2 - 1
quant = dtype in (torch.qint8, torch.quint8)
if quant:
    x = torch.quantize_per_tensor(torch.rand(shape, device=dev), 1 / 255., -128, dtype)
    wq = quantize_shift_weights(w * 3)
    if sys.argv[1].endswith("cl"):
        x = x.contiguous(memory_format=torch.channels_last)
else:
    x = torch.randn(shape, device=dev).to(dtype); g = torch.randn(shape, device=dev).to(dtype); w = w.to(dtype)
n = x.numel()
es = 1 if quant else x.element_size()


def graph_time(fn, reps=30):
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
    for _ in range(3):
        gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


for spec in specs:
    if spec and lib.ts_set_tuning(spec.encode()) != 0:
        print(spec, "rejected"); continue
    try:
        if quant:
            tf = graph_time(lambda: fwd(x, wq, borders, list(shape), pad, False)); tb = float("nan")
        else:
            tf = graph_time(lambda: fwd(x, w, borders, list(shape), pad, active))
            tb = graph_time(lambda: bwd(g, w, x, borders, pad, active))
    except RuntimeError as e:
        print(spec, "failed", str(e)[:100]); continue
    print(f"{sys.argv[1]} [{spec or 'defaults'}]: fwd {tf:7.1f} us ({n * 2 * es / tf / 6541.8e3:4.0%})  bwd {tb:7.1f} us ({n * 3 * es / tb / 6541.8e3:4.0%})", flush=True)
