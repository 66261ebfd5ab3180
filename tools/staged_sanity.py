#!/usr/bin/env python
"""Small staged-path (bulk-async kernels) parity run, meant to be executed under compute-sanitizer
on the GPU box:   compute-sanitizer --tool memcheck python tools/staged_sanity.py
Every case forces one kernel family (--path=2 staged (default), 3 TMA, 5 halo, 1 generic) and compares with the CPU oracle;
failures are listed, not raised one by one, so a single GPU call reports everything."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]

import torchshifts  # noqa: E402
from torchshifts.extension import native  # noqa: E402
from torchshifts import functional as F  # noqa: E402
from oracle.oracle import Oracle, quantize_shift_weights_np  # noqa: E402

lib = native().lib
orc = Oracle("port")
dev = torch.device("cuda:0")
fn = {1: F.shift1d_func, 2: F.shift2d_func, 3: F.shift3d_func}
fails, ran, skipped = [], 0, 0
quick = "--quick" in sys.argv
TUNING = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tuning=")), None)
if TUNING:
    assert lib.ts_set_tuning(TUNING.encode()) == 0, TUNING
FORCED = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--path=")), 2)

shapes = [(3, 5, 64), (2, 4, 40), (2, 3, 12, 16), (3, 4, 8, 8), (2, 6, 28, 28), (2, 3, 4, 6, 8), (1, 2, 3, 5, 4), (5, 7, 1, 16), (9, 2, 4, 12), (37, 3, 8, 8),
          (2, 3, 7, 6, 16), (1, 2, 1, 6, 8), (2, 2, 5, 1, 8), (3, 2, 1040), (3, 5, 5, 9, 12), (2, 3, 9, 20), (5, 2, 3, 7, 8)]
if quick:
    shapes = shapes[:5]
rng = np.random.default_rng(0)
lib.ts_set_kernel_path(FORCED)
for shape in shapes:
    dim = len(shape) - 2
    for wr in (2.5, 40.0):
        x = rng.standard_normal(shape).astype(np.float32)
        w = ((rng.random((shape[1], dim)) * 2 - 1) * wr).astype(np.float32)
        seen_b = set()
        for borders in (None, [[0, 4]] * dim if min(shape[2:]) > 4 else None, [[1, 2], [2, 1], [1, 1]][:dim] if min(shape[2:]) > 5 else None,
                        [[0, 0]] * (dim - 1) + [[4, 0]] if shape[-1] > 8 else None):
            if str(borders) in seen_b:
                continue
            seen_b.add(str(borders))
            for pad in range(5):
                for active in (False, True):
                    tag = (shape, wr, borders, pad, active)
                    y_ref = orc.forward(x, w, pad, active, borders)
                    g = rng.standard_normal(y_ref.shape).astype(np.float32)
                    xd = torch.from_numpy(x).to(dev).requires_grad_(True)
                    wd = torch.from_numpy(w).to(dev).requires_grad_(True)
                    b = torch.tensor(borders, dtype=torch.long) if borders else None
                    try:
                        y = fn[dim](xd, wd, pad, active, b)
                        torch.cuda.synchronize()
                    except RuntimeError as e:
                        if "UNSUPPORTED" in str(e):
                            skipped += 1
                        else:
                            fails.append((tag, "fwd exception " + str(e)[:200]))
                        y = None
                    if y is not None:
                        ran += 1
                        if not np.array_equal(y.detach().cpu().numpy(), y_ref):
                            fails.append((tag, "forward", int((y.detach().cpu().numpy() != y_ref).sum())))
                    if y is None:      # forward family not applicable: still exercise backward through the auto path
                        lib.ts_set_kernel_path(0)
                        y = fn[dim](xd, wd, pad, active, b)
                        lib.ts_set_kernel_path(FORCED)
                    try:
                        y.backward(torch.from_numpy(g).to(dev))
                        torch.cuda.synchronize()
                    except RuntimeError as e:
                        if "UNSUPPORTED" in str(e):
                            skipped += 1
                        else:
                            fails.append((tag, "bwd exception " + str(e)[:200]))
                        continue
                    ran += 1
                    gi_ref, _ = orc.backward(g, x, w, pad, active, borders)
                    _, gw64 = orc.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active, borders)
                    if not np.array_equal(xd.grad.cpu().numpy(), gi_ref):
                        fails.append((tag, "grad_input", int((xd.grad.cpu().numpy() != gi_ref).sum())))
                    if not np.allclose(wd.grad.cpu().numpy(), gw64, rtol=1e-5, atol=1e-5 * np.abs(gw64).max() + 1e-30):
                        fails.append((tag, "grad_weight", float(np.abs(wd.grad.cpu().numpy() - gw64).max())))
# raw element sizes through the sparse gather: f64, f16, int8 (quantized), int32 (quantized)
for shape in [(2, 3, 8, 16), (3, 2, 64), (2, 2, 4, 4, 8), (4, 3, 6, 24), (3, 4, 56, 56), (2, 2, 10, 40), (2, 3, 7, 24), (5, 2, 2, 24), (2, 2, 6, 20)]:
    dim = len(shape) - 2
    w = ((rng.random((shape[1], dim)) * 2 - 1) * 3).astype(np.float32)
    for pad in range(5):
        for tdt in (torch.float64, torch.float16, torch.bfloat16):
            xt = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(tdt)
            try:
                y = fn[dim](xt.to(dev), torch.from_numpy(w).to(tdt).to(dev), pad, False)
            except RuntimeError as e:
                skipped += 1 if "UNSUPPORTED" in str(e) else 0
                if "UNSUPPORTED" not in str(e):
                    fails.append(((shape, pad, str(tdt)), "exception " + str(e)[:200]))
                continue
            ran += 1
            want = orc.forward(xt.double().numpy(), torch.from_numpy(w).to(tdt).double().numpy(), pad, False)
            if not np.array_equal(y.double().cpu().numpy(), want):
                fails.append(((shape, pad, str(tdt)), "gather", int((y.double().cpu().numpy() != want).sum())))
        raw, wzp = quantize_shift_weights_np(w)
        for qdt, zp in ((torch.quint8, 3), (torch.qint8, -128), (torch.qint32, 11)):
            xq = torch.quantize_per_tensor(torch.rand(shape).to(dev), 0.01, zp, qdt)
            qw = torch._make_per_tensor_quantized_tensor(torch.from_numpy(raw.astype(np.uint8)).to(dev), 1.0, wzp)
            try:
                yq = torch.ops.torchshifts.__getattr__(f"shift{dim}d")(xq, qw, torch.Tensor(), pad, False)
            except RuntimeError as e:
                skipped += 1 if "UNSUPPORTED" in str(e) else 0
                if "UNSUPPORTED" not in str(e):
                    fails.append(((shape, pad, str(qdt)), "exception " + str(e)[:200]))
                continue
            ran += 1
            want = orc.qforward(xq.int_repr().cpu().numpy(), raw, wzp, zp, pad)
            if not np.array_equal(yq.int_repr().cpu().numpy(), want):
                fails.append(((shape, pad, str(qdt)), "qgather", int((yq.int_repr().cpu().numpy() != want).sum())))
lib.ts_set_kernel_path(0)
torch.cuda.synchronize()
print(f"sanity (forced path {FORCED}): ran {ran}, skipped (staged path not applicable) {skipped}, failures {len(fails)}")
for f in fails[:60]:
    print("  FAIL", f)
sys.exit(1 if fails else 0)
