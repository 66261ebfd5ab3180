import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts
from torchshifts.extension import native
from torchshifts.quantized.functional import shift2d_quantized
from torchshifts.quantized.modules.shifts import quantize_shift_weights
lib = native().lib
dev = torch.device("cuda:0")
N, spread, variant = int(sys.argv[1]), float(sys.argv[2]), sys.argv[3]
torch.manual_seed(5)
x = torch.rand(N, 256, 56, 56, device=dev)
w = (torch.rand(256, 2, device=dev) * 2 - 1) * spread
qw = quantize_shift_weights(w)
xq = torch.quantize_per_tensor(x, 1 / 255., -128, torch.qint8)
xcl = xq.contiguous(memory_format=torch.channels_last)
planar = shift2d_quantized(xq, qw, 0).int_repr()
assert lib.ts_set_tuning(variant.encode()) == 0
torch.cuda.synchronize()
t0 = time.time()
y = shift2d_quantized(xcl, qw, 0)
torch.cuda.synchronize()
print(f"N={N} spread={spread} {variant}: first call {1e3 * (time.time() - t0):.2f} ms, equal={torch.equal(y.int_repr(), planar)}", flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    y = shift2d_quantized(xcl, qw, 0)
b.record(); torch.cuda.synchronize()
print(f"   {a.elapsed_time(b) / 10 * 1000:.1f} us per call", flush=True)
