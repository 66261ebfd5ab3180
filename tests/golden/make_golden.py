#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own torch extension.

Run once in the build container (where /root/reference exists):

    ./oracle/build_ref_full.sh          # -> oracle/_ref/torchshifts_ref/_C.so  (unmodified csrc)
    python tests/golden/make_golden.py  # -> tests/golden/shift_golden.npz, quant_golden.npz, kat.npz

The script loads ONLY the reference library (torch.ops.load_library) and calls its registered
ops ``torchshifts::shift{1,2,3}d`` (csrc/torchshifts.cpp:35-40) and their autograd
(csrc/ops/autograd/shifts_autograd.cpp) with ``torch.set_num_threads(1)`` so that grad_weight is
the deterministic serial sum.  It must not import this repository's ``torchshifts``
package: both register the ``torchshifts::`` namespace.  The fixtures travel with the repo; the
GPU box never needs /root/reference.
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
LIB = ROOT / "oracle" / "_ref" / "torchshifts_ref" / "_C.so"
OUT = Path(__file__).resolve().parent

SHAPES = {1: (2, 3, 9), 2: (2, 3, 5, 6), 3: (2, 2, 4, 5, 3)}
BORDERS = {1: [[1, 2]], 2: [[1, 1], [2, 1]], 3: [[1, 0], [0, 2], [1, 1]]}


def weights_for(dim, C, gen, dtype):
    # mix of fractional shifts, exact .5 ties, zero, and shifts larger than the axis length
    base = (torch.rand(C, dim, generator=gen, dtype=torch.float64) * 2 - 1) * 3.7
    special = torch.tensor([0.5, -0.5, 1.5, -1.5, 2.5, 0.0, 7.25, -11.0, 4.0], dtype=torch.float64)
    flat = base.flatten()
    k = min(flat.numel(), special.numel())
    idx = torch.randperm(flat.numel(), generator=gen)[:k]
    flat[idx] = special[torch.randperm(special.numel(), generator=gen)[:k]]
    return flat.reshape(C, dim).to(dtype)


def main():
    if not LIB.exists():
        sys.exit(f"{LIB} not found: run oracle/build_ref_full.sh first")
    torch.ops.load_library(str(LIB))
    torch.set_num_threads(1)
    ops = {1: torch.ops.torchshifts.shift1d, 2: torch.ops.torchshifts.shift2d, 3: torch.ops.torchshifts.shift3d}
    gen = torch.Generator().manual_seed(20261017)

    # ---------------------------------------------------------------- float forward/backward
    store = {}
    names = []
    for dtype, dname in ((torch.float32, "f32"), (torch.float64, "f64")):
        for dim, shape in SHAPES.items():
            for trial in range(2):
                x0 = torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
                w0 = weights_for(dim, shape[1], gen, dtype)
                for use_b in (0, 1):
                    borders = torch.tensor(BORDERS[dim], dtype=torch.long) if use_b else torch.Tensor()
                    for pad in range(5):
                        for active in (0, 1):
                            x = x0.clone().requires_grad_(True)
                            w = w0.clone().requires_grad_(True)
                            y = ops[dim](x, w, borders, pad, bool(active))
                            g = torch.randn(y.shape, generator=gen, dtype=torch.float64).to(dtype)
                            y.backward(g)
                            name = f"{dname}_d{dim}_t{trial}_b{use_b}_p{pad}_a{active}"
                            names.append(name)
                            store[name + "/x"] = x0.numpy()
                            store[name + "/w"] = w0.numpy()
                            store[name + "/g"] = g.numpy()
                            store[name + "/y"] = y.detach().numpy()
                            store[name + "/gi"] = x.grad.numpy()
                            store[name + "/gw"] = w.grad.numpy()
    # channels-last input (reference nhwdc bodies, kernels/shifts_kernels.h:330-527)
    for dim, shape, fmt in ((2, SHAPES[2], torch.channels_last), (3, SHAPES[3], torch.channels_last_3d)):
        x0 = torch.randn(shape, generator=gen).contiguous(memory_format=fmt)
        w0 = weights_for(dim, shape[1], gen, torch.float32)
        for pad in (0, 3):
            for active in (0, 1):
                x = x0.clone(memory_format=torch.preserve_format).requires_grad_(True)
                assert x.is_contiguous(memory_format=fmt)
                w = w0.clone().requires_grad_(True)
                y = ops[dim](x, w, torch.Tensor(), pad, bool(active))
                g = torch.randn(y.shape, generator=gen)
                y.backward(g)
                name = f"cl_d{dim}_p{pad}_a{active}"
                names.append(name)
                store[name + "/x"] = x0.contiguous().numpy()   # logical NCHW values
                store[name + "/w"] = w0.numpy()
                store[name + "/g"] = g.numpy()
                store[name + "/y"] = y.detach().contiguous().numpy()
                store[name + "/gi"] = x.grad.contiguous().numpy()
                store[name + "/gw"] = w.grad.numpy()
    store["names"] = np.array(names)
    np.savez_compressed(OUT / "shift_golden.npz", **store)
    print(f"shift_golden.npz: {len(names)} cases")

    # ---------------------------------------------------------------- quantized forward
    qstore, qnames = {}, []

    def quantize_shift_weights(weight):  # torchshifts/quantized/modules/shifts.py:10-12
        scale = math.ceil((weight.max().item() - weight.min().item()) / 255.)
        return torch.quantize_per_tensor(weight, scale, 128, torch.quint8)

    qcfg = ((torch.quint8, 1 / 255., 0, "quint8"), (torch.qint8, 1 / 255., -128, "qint8"),
            (torch.quint8, 0.02, 37, "quint8zp"), (torch.qint32, 1e-3, 11, "qint32"))
    for dim, shape in SHAPES.items():
        xf = torch.rand(shape, generator=gen)
        for wscale, wtag in ((3.0, "w3"), (200.0, "w200")):
            wf = (torch.rand(shape[1], dim, generator=gen) * 2 - 1) * wscale
            qw = quantize_shift_weights(wf)
            for qdtype, scale, zp, qtag in qcfg:
                xq = torch.quantize_per_tensor(xf, scale, zp, qdtype)
                for use_b in (0, 1):
                    borders = torch.tensor(BORDERS[dim], dtype=torch.long) if use_b else torch.Tensor()
                    for pad in range(5):
                        yq = ops[dim](xq, qw, borders, pad, False)
                        assert yq.q_scale() == xq.q_scale() and yq.q_zero_point() == xq.q_zero_point()
                        assert yq.dtype == xq.dtype
                        name = f"{qtag}_d{dim}_{wtag}_b{use_b}_p{pad}"
                        qnames.append(name)
                        qstore[name + "/x"] = xq.int_repr().numpy()
                        qstore[name + "/wf"] = wf.numpy()
                        qstore[name + "/wq"] = qw.int_repr().numpy()
                        qstore[name + "/meta"] = np.array([qw.q_zero_point(), zp], dtype=np.int64)
                        qstore[name + "/wscale"] = np.array([qw.q_scale()], dtype=np.float64)
                        qstore[name + "/xscale"] = np.array([scale], dtype=np.float64)
                        qstore[name + "/y"] = yq.int_repr().numpy()
    qstore["names"] = np.array(qnames)
    np.savez_compressed(OUT / "quant_golden.npz", **qstore)
    print(f"quant_golden.npz: {len(qnames)} cases")

    # ---------------------------------------------------------------- known-answer table
    kat = {}
    x = torch.tensor([10., 11., 12., 13., 14.]).reshape(1, 1, 5)
    shifts = [-7, -4, -1, 0, 1, 2, 4, 5, 7]
    table = np.zeros((len(shifts), 5, 5), dtype=np.float32)
    for si, s in enumerate(shifts):
        for pad in range(5):
            table[si, pad] = ops[1](x, torch.tensor([[float(s)]]), torch.Tensor(), pad, False).numpy().reshape(5)
    kat["shift1d_shifts"] = np.array(shifts)
    kat["shift1d_table"] = table
    x2 = torch.arange(120, dtype=torch.float32).reshape(2, 3, 4, 5)
    w2 = torch.tensor([[1., 0.], [0., -1.], [0.4, 1.6]])
    kat["shift2d_arange_y"] = ops[2](x2, w2, torch.Tensor(), 0, False).numpy()
    ties = torch.tensor([[0.5], [1.5], [2.5], [-0.5], [-1.5], [-2.5], [3.5]])
    xt = torch.arange(7 * 6, dtype=torch.float32).reshape(1, 7, 6)
    kat["ties_w"] = ties.numpy()
    kat["ties_y"] = ops[1](xt, ties, torch.Tensor(), 2, False).numpy()
    kat["size1_reflect_y"] = ops[2](torch.arange(4.).reshape(1, 1, 1, 4), torch.tensor([[3., 1.]]),
                                    torch.Tensor(), 3, False).numpy()
    kat["len2_reflect_y"] = ops[2](torch.arange(8.).reshape(1, 1, 2, 4), torch.tensor([[1., 0.]]),
                                   torch.Tensor(), 3, False).numpy()
    # border validation corner cases (csrc/ops/shifts.cpp:93-135): output shapes only
    bshapes = []
    xb = torch.zeros(1, 1, 16, 16)
    for b in ([[16, 0], [3, 13]], [[0, 0], [0, 0]], [[15, 0], [0, 15]], [[3, 4], [5, 6]], [[0, 16], [8, 8]],
              [[10, 10], [1, 1]]):
        yb = ops[2](xb, torch.zeros(1, 2), torch.tensor(b, dtype=torch.long), 0, False)
        bshapes.append(b[0] + b[1] + list(yb.shape[2:]))
    kat["border_cases"] = np.array(bshapes, dtype=np.int64)
    np.savez_compressed(OUT / "kat.npz", **kat)
    print("kat.npz written")


if __name__ == "__main__":
    main()
