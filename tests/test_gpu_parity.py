"""GPU parity: the CUDA path (public API -> torch.ops.torchshifts -> C ABI -> sm_100a kernels)
against the CPU oracle on the same seeded inputs, against the golden fixtures produced by the
reference's own extension, and -- at BASELINE.json's full sizes -- through size-independent
properties plus an oracle check of sampled images.

Tolerances (BASELINE.json north_star):
  * forward (sparse, active, quantized) and grad_input: BIT-EXACT for fp32 / fp64 / integer;
  * grad_weight: rtol 1e-5 against the fp64 oracle, atol 1e-5 * max|gw| (the fp32 oracle's own
    serial sum is less accurate than that);
  * fp16 / bf16 (no CPU oracle exists): fp32 oracle on the rounded inputs, rtol 1e-2.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.gpu

BORDERS = {1: [[1, 2]], 2: [[1, 1], [2, 1]], 3: [[1, 0], [0, 2], [1, 1]]}
GENERIC, STAGED, TMA, HALO, FLAT = 1, 2, 3, 5, 6
BANDWIDTH = (STAGED, TMA, HALO, FLAT)        # the shared-memory staged families (anything but the generic kernels)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    from torchshifts.extension import native
    return native().lib


@pytest.fixture()
def auto_path(lib):
    lib.ts_set_kernel_path(0)
    lib.ts_set_tuning(b"use_tma=1,use_halo=1,use_flat=1,nhwc_variant=0,nhwc_ring_rows=0")
    yield
    lib.ts_set_kernel_path(0)
    lib.ts_set_tuning(b"use_tma=1,use_halo=1,use_flat=1,nhwc_variant=0,nhwc_ring_rows=0")


# kernel-family selection modes exercised by the sweeps: the automatic choice (TMA-tensor family for
# zeros padding, bulk-staged otherwise, generic for odd shapes), the automatic choice without the
# TMA-tensor family (so the bulk-staged kernels also see zeros padding), and the generic family only.
# "staged" switches the round-2 families (halo, flat) off as well, so the bulk-staged kernels keep their full coverage.
MODES = ["auto", "no_tma", "staged", "generic"]


def _set_mode(lib, mode):
    lib.ts_set_kernel_path(GENERIC if mode == "generic" else 0)
    lib.ts_set_tuning({"no_tma": b"use_tma=0,use_halo=1,use_flat=1", "staged": b"use_tma=0,use_halo=0,use_flat=0"}.get(
        mode, b"use_tma=1,use_halo=1,use_flat=1"))


def _func(dim):
    from torchshifts import functional as F
    return {1: F.shift1d_func, 2: F.shift2d_func, 3: F.shift3d_func}[dim]


def _run_cuda(dev, dim, x, w, g, pad, active, borders):
    xd = torch.from_numpy(x).to(dev).requires_grad_(True)
    wd = torch.from_numpy(w).to(dev).requires_grad_(True)
    b = torch.tensor(borders, dtype=torch.long) if borders is not None else None
    y = _func(dim)(xd, wd, pad, active, b)
    if g is None:
        return y.detach().cpu().numpy(), None, None
    y.backward(torch.from_numpy(g).to(dev))
    return y.detach().cpu().numpy(), xd.grad.cpu().numpy(), wd.grad.cpu().numpy()


def _gw_close(gw, gw64):
    scale = np.abs(gw64).max()
    return np.allclose(gw.astype(np.float64), gw64, rtol=1e-5, atol=1e-5 * scale + 1e-30)


def _case_args(name):
    parts = name.split("_")
    if parts[0] == "cl":
        return int(parts[1][1]), int(parts[2][1]), bool(int(parts[3][1])), None
    return int(parts[1][1]), int(parts[4][1]), bool(int(parts[5][1])), (BORDERS[int(parts[1][1])] if int(parts[3][1]) else None)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_golden_fixtures_float(dev, lib, golden, oracle_port, mode):
    """All 248 reference-generated cases (3 dims x 5 paddings x sparse/active x borders x fp32/fp64)."""
    _set_mode(lib, mode)
    try:
        g = golden["shift_golden"]
        for name in g["names"].tolist():
            dim, pad, active, borders = _case_args(name)
            x, w, grad = g[name + "/x"], g[name + "/w"], g[name + "/g"]
            y, gi, gw = _run_cuda(dev, dim, x, w, grad, pad, active, borders)
            assert np.array_equal(y, g[name + "/y"]), f"forward {name}"
            assert np.array_equal(gi, g[name + "/gi"]), f"grad_input {name}"
            _, gw64 = oracle_port.backward(grad.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active, borders)
            assert _gw_close(gw, gw64), f"grad_weight {name}: {gw} vs {gw64}"
            assert np.allclose(gw, g[name + "/gw"], rtol=1e-4, atol=1e-4 * np.abs(gw64).max()), f"grad_weight vs reference fp sum {name}"
    finally:
        _set_mode(lib, "auto")


def test_golden_fixtures_quantized(dev, lib, golden, auto_path):
    from torchshifts.quantized.functional import shift1d_quantized, shift2d_quantized, shift3d_quantized
    fns = {1: shift1d_quantized, 2: shift2d_quantized, 3: shift3d_quantized}
    qd = {"quint8": torch.quint8, "qint8": torch.qint8, "quint8zp": torch.quint8, "qint32": torch.qint32}
    g = golden["quant_golden"]
    for mode in MODES:
        _set_mode(lib, mode)
        for name in g["names"].tolist():
            parts = name.split("_")
            dim, use_b, pad = int(parts[1][1]), int(parts[3][1]), int(parts[4][1])
            wzp, zp = g[name + "/meta"].tolist()
            xq = torch._make_per_tensor_quantized_tensor(torch.from_numpy(g[name + "/x"]).to(dev), float(g[name + "/xscale"][0]), int(zp))
            assert xq.dtype == qd[parts[0]]
            qw = torch._make_per_tensor_quantized_tensor(torch.from_numpy(g[name + "/wq"]).to(dev), float(g[name + "/wscale"][0]) or 1.0, int(wzp))
            b = torch.tensor(BORDERS[dim], dtype=torch.long) if use_b else None
            yq = fns[dim](xq, qw, pad, b)
            assert yq.is_quantized and yq.dtype == xq.dtype
            assert yq.q_scale() == xq.q_scale() and yq.q_zero_point() == xq.q_zero_point()
            assert np.array_equal(yq.int_repr().cpu().numpy(), g[name + "/y"]), name


def test_known_answers(dev, golden, auto_path):
    kat = golden["kat"]
    x = np.array([10, 11, 12, 13, 14], dtype=np.float32).reshape(1, 1, 5)
    for si, s in enumerate(kat["shift1d_shifts"].tolist()):
        for pad in range(5):
            y, _, _ = _run_cuda(dev, 1, x, np.array([[s]], dtype=np.float32), None, pad, False, None)
            assert np.array_equal(y.reshape(-1), kat["shift1d_table"][si, pad]), (s, pad)
    x2 = np.arange(120, dtype=np.float32).reshape(2, 3, 4, 5)
    y, _, _ = _run_cuda(dev, 2, x2, np.array([[1., 0.], [0., -1.], [0.4, 1.6]], dtype=np.float32), None, 0, False, None)
    assert np.array_equal(y, kat["shift2d_arange_y"])
    xt = np.arange(42, dtype=np.float32).reshape(1, 7, 6)
    y, _, _ = _run_cuda(dev, 1, xt, kat["ties_w"].astype(np.float32), None, 2, False, None)
    assert np.array_equal(y, kat["ties_y"])          # half-to-even, like the CPU reference
    y, _, _ = _run_cuda(dev, 2, np.arange(4, dtype=np.float32).reshape(1, 1, 1, 4), np.array([[3., 1.]], dtype=np.float32), None, 3, False, None)
    assert np.array_equal(y, kat["size1_reflect_y"])
    y, _, _ = _run_cuda(dev, 2, np.arange(8, dtype=np.float32).reshape(1, 1, 2, 4), np.array([[1., 0.]], dtype=np.float32), None, 3, False, None)
    assert np.array_equal(y, kat["len2_reflect_y"])


SWEEP = [
    # (shape, weight range) -- sizes picked so that the staged path applies to most (rows % 4 == 0)
    ((3, 5, 64), 9.0), ((2, 4, 40), 70.0), ((2, 3, 12, 16), 3.0), ((3, 4, 8, 8), 12.0), ((2, 6, 28, 28), 2.0),
    ((1, 3, 7, 9), 3.0), ((2, 2, 5, 8), 4.0), ((2, 3, 4, 6, 8), 2.5), ((1, 2, 3, 5, 4), 5.0), ((5, 7, 1, 16), 3.0),
    ((2, 3, 16, 1), 3.0), ((1, 1, 2, 2), 1.5), ((9, 2, 4, 12), 2.0),
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_randomised_sweep_vs_oracle(dev, lib, oracle_port, mode, dtype):
    _set_mode(lib, mode)
    staged_hits = tma_hits = 0
    try:
        rng = np.random.default_rng(123)
        for shape, wr in SWEEP:
            dim = len(shape) - 2
            x = rng.standard_normal(shape).astype(dtype)
            w = ((rng.random((shape[1], dim)) * 2 - 1) * wr).astype(dtype)
            w.flat[0] = 0.5
            for borders in (None, [[1, 0]] * dim if min(shape[2:]) > 1 else None, [[0, 4]] * dim if min(shape[2:]) > 4 else None):
                for pad in range(5):
                    for active in (False, True):
                        y_ref = oracle_port.forward(x, w, pad, active, borders)
                        grad = rng.standard_normal(y_ref.shape).astype(dtype)
                        y, gi, gw = _run_cuda(dev, dim, x, w, grad, pad, active, borders)
                        staged_hits += lib.ts_last_kernel_path() in BANDWIDTH
                        tma_hits += lib.ts_last_kernel_path() == TMA
                        tag = (shape, pad, active, borders, dtype.__name__)
                        assert np.array_equal(y, y_ref), ("forward",) + tag
                        gi_ref, _ = oracle_port.backward(grad, x, w, pad, active, borders)
                        assert np.array_equal(gi, gi_ref), ("grad_input",) + tag
                        _, gw64 = oracle_port.backward(grad.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active, borders)
                        assert _gw_close(gw, gw64), ("grad_weight",) + tag + (gw, gw64)
    finally:
        _set_mode(lib, "auto")
    if mode == "generic":
        assert staged_hits == 0
    if mode in ("no_tma", "staged"):
        assert tma_hits == 0
    if mode == "auto" and dtype is np.float32:
        assert tma_hits > 0 and staged_hits > tma_hits


def test_half_precision_vs_fp32_oracle(dev, oracle_port, auto_path):
    rng = np.random.default_rng(5)
    for tdtype in (torch.float16, torch.bfloat16):
        for shape in ((2, 4, 64), (2, 3, 8, 16), (1, 2, 4, 4, 8)):
            dim = len(shape) - 2
            x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(tdtype)
            w = torch.from_numpy(((rng.random((shape[1], dim)) * 2 - 1) * 3).astype(np.float32)).to(tdtype)
            g = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(tdtype)
            for pad in range(5):
                for active in (False, True):
                    xd, wd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
                    y = _func(dim)(xd, wd, pad, active)
                    y.backward(g.to(dev))
                    x32, w32, g32 = x.float().numpy(), w.float().numpy(), g.float().numpy()
                    y_ref = oracle_port.forward(x32, w32, pad, active)
                    gi_ref, gw_ref = oracle_port.backward(g32, x32, w32, pad, active)
                    if not active:
                        assert np.array_equal(y.detach().float().cpu().numpy(), y_ref)          # pure copy: exact
                        assert np.array_equal(xd.grad.float().cpu().numpy(), gi_ref)
                    else:
                        assert np.allclose(y.detach().float().cpu().numpy(), y_ref, rtol=1e-2, atol=1e-2)
                        assert np.allclose(xd.grad.float().cpu().numpy(), gi_ref, rtol=1e-2, atol=1e-2)
                    assert np.allclose(wd.grad.float().cpu().numpy(), gw_ref, rtol=1e-2, atol=1e-2 * np.abs(gw_ref).max())


def test_rows_every_window_phase(dev, oracle_port, auto_path):
    """1-D tensors (cfg2's layout) through the flat row loop of the staged family: shifts in +-9 put the x and grad
    windows on every byte phase of a 16-byte group (8 phases for 16-bit elements), crops untie the grad_input window's
    phase from the x window's (run-time realignment) or leave no aligned interior at all.  16-bit results must equal the
    fp32 oracle rounded ONCE to the storage type (fp32 arithmetic, one rounding at the store)."""
    rng = np.random.default_rng(17)
    from torchshifts.functional import shift1d_func
    for tdtype in (torch.bfloat16, torch.float16, torch.float32):
        for shape in ((3, 20, 256), (2, 9, 1040)):
            x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(tdtype)
            w = torch.from_numpy(((rng.random((shape[1], 1)) * 2 - 1) * 9).astype(np.float32)).to(tdtype)
            x32, w32 = x.float().numpy(), w.float().numpy()
            for borders in (None, [[8, 0]], [[16, 24]], [[3, 2]]):
                b = torch.tensor(borders, dtype=torch.long) if borders else None
                for pad in range(5):
                    for active in (False, True):
                        y_ref = oracle_port.forward(x32, w32, pad, active, borders)
                        g = torch.from_numpy(rng.standard_normal(y_ref.shape).astype(np.float32)).to(tdtype)
                        g32 = g.float().numpy()
                        xd, wd = x.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
                        y = shift1d_func(xd, wd, pad, active, b)
                        y.backward(g.to(dev))
                        gi_ref, _ = oracle_port.backward(g32, x32, w32, pad, active, borders)
                        _, gw64 = oracle_port.backward(g32.astype(np.float64), x32.astype(np.float64), w32.astype(np.float64), pad, active, borders)
                        tag = (str(tdtype), shape, borders, pad, active)
                        assert torch.equal(y.detach().cpu(), torch.from_numpy(y_ref).to(tdtype)), ("forward",) + tag
                        assert torch.equal(xd.grad.cpu(), torch.from_numpy(gi_ref).to(tdtype)), ("grad_input",) + tag
                        tol = 1e-5 if tdtype is torch.float32 else 1e-2
                        assert np.allclose(wd.grad.float().cpu().numpy(), gw64, rtol=tol, atol=tol * np.abs(gw64).max() + 1e-30), ("grad_weight",) + tag


def test_more_channels_than_the_shift_table_holds(dev, lib, oracle_port, auto_path):
    """The bandwidth families tabulate the per-channel shift parameters in shared memory for C <= 512 and (halo family) sort
    the channels by window-misalignment class there; above that they fall back to per-unit evaluation / a single run-time
    class.  C = 520 through every family that accepts the shape (forced), all paddings, against the oracle."""
    rng = np.random.default_rng(23)
    for shape, dim in (((2, 520, 8, 16), 2), ((1, 520, 3, 6, 8), 3), ((2, 520, 64), 1)):
        x = rng.standard_normal(shape).astype(np.float32)
        w = ((rng.random((shape[1], dim)) * 2 - 1) * 2.5).astype(np.float32)
        for pad in range(5):
            for active in (False, True):
                y_ref = oracle_port.forward(x, w, pad, active)
                g = rng.standard_normal(y_ref.shape).astype(np.float32)
                gi_ref, _ = oracle_port.backward(g, x, w, pad, active)
                _, gw64 = oracle_port.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active)
                for path in (0, 2, 3, 5):
                    lib.ts_set_kernel_path(path)
                    try:
                        y, gi, gw = _run_cuda(dev, dim, x, w, g, pad, active, None)
                    except RuntimeError as e:
                        assert path != 0 and "UNSUPPORTED" in str(e), (shape, pad, active, path, str(e)[:120])
                        continue
                    finally:
                        lib.ts_set_kernel_path(0)
                    tag = (shape, pad, active, path)
                    assert np.array_equal(y, y_ref), ("forward",) + tag
                    assert np.array_equal(gi, gi_ref), ("grad_input",) + tag
                    assert _gw_close(gw, gw64), ("grad_weight",) + tag


def test_host_pipeline_matches_the_operators(dev, oracle_port, auto_path):
    """torchshifts.host.HostShiftPipeline (bench.py's e2e leg: pinned host buffers, three streams, chunks through a ring of
    device slots, operator outputs handed back to the allocator only behind the copy-out) gives what one call of the
    operators on the whole tensor gives, step after step, also with a ragged last chunk and more chunks than slots."""
    from torchshifts.host import HostShift2dPipeline
    from torchshifts.functional import shift2d_func
    torch.manual_seed(11)
    N, C, H, W = 23, 8, 12, 16
    pipe = HostShift2dPipeline(N, C, H, W, device=dev, chunk=4, slots=3)
    w = (torch.rand(C, 2, device=dev) * 2 - 1) * 2
    for step, (pad, active) in enumerate(((0, False), (3, True), (0, False))):
        pipe.x_host.normal_(); pipe.g_host.normal_()
        gw = pipe.forward_backward(w, pad, active)
        pipe.read_back_grad_weight(gw)
        torch.cuda.synchronize()
        xd = pipe.x_host.to(dev).requires_grad_(True)
        wd = w.clone().requires_grad_(True)
        y = shift2d_func(xd, wd, pad, active)
        y.backward(pipe.g_host.to(dev))
        assert torch.equal(pipe.y_host, y.detach().cpu()), step
        assert torch.equal(pipe.gi_host, xd.grad.cpu()), step
        assert torch.allclose(pipe.gw_host, wd.grad.cpu(), rtol=1e-5, atol=1e-5 * float(wd.grad.abs().max())), step
        y_ref = oracle_port.forward(pipe.x_host.numpy(), w.cpu().numpy(), pad, active)
        assert np.array_equal(pipe.y_host.numpy(), y_ref), step


def test_strided_and_channels_last_inputs(dev, oracle_port, auto_path):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 6, 8, 12)).astype(np.float32)
    w = ((rng.random((6, 2)) * 2 - 1) * 3).astype(np.float32)
    g = rng.standard_normal(x.shape).astype(np.float32)
    from torchshifts.functional import shift2d_func
    for pad, active in ((0, False), (3, True), (2, True)):
        y_ref = oracle_port.forward(x, w, pad, active)
        gi_ref, gw_ref = oracle_port.backward(g, x, w, pad, active)
        for make in (lambda t: t.contiguous(memory_format=torch.channels_last),
                     lambda t: torch.cat([t, t], dim=3)[..., :12],                 # row stride 24
                     lambda t: t.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)):
            xd = make(torch.from_numpy(x).to(dev)).requires_grad_(True)
            wd = torch.from_numpy(w).to(dev).requires_grad_(True)
            y = shift2d_func(xd, wd, pad, active)
            assert y.is_contiguous()
            y.backward(torch.from_numpy(g).to(dev))
            assert np.array_equal(y.detach().cpu().numpy(), y_ref)
            assert np.array_equal(xd.grad.cpu().numpy(), gi_ref)
            assert np.allclose(wd.grad.cpu().numpy(), gw_ref, rtol=1e-4, atol=1e-4)


def test_large_channels_last_input_takes_the_bandwidth_path(dev, lib, oracle_port, auto_path):
    """Above the copy threshold a channels-last input is made dense once and served by the TMA / staged
    kernels (values identical to the strided generic path and to the oracle)."""
    from torchshifts.functional import shift2d_func
    rng = np.random.default_rng(21)
    x = rng.standard_normal((4, 32, 24, 32)).astype(np.float32)       # 98 304 elements
    w = ((rng.random((32, 2)) * 2 - 1) * 2).astype(np.float32)
    g = rng.standard_normal(x.shape).astype(np.float32)
    for pad, active in ((0, False), (4, True)):
        xd = torch.from_numpy(x).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        wd = torch.from_numpy(w).to(dev).requires_grad_(True)
        y = shift2d_func(xd, wd, pad, active)
        assert lib.ts_last_kernel_path() in BANDWIDTH
        before = lib.ts_launch_count()
        y.backward(torch.from_numpy(g).to(dev))
        assert lib.ts_last_kernel_path() in BANDWIDTH
        assert lib.ts_launch_count() - before == 2, "the backward must reuse the forward's planar copy (no second layout pass)"
        assert np.array_equal(y.detach().cpu().numpy(), oracle_port.forward(x, w, pad, active))
        gi_ref, _ = oracle_port.backward(g, x, w, pad, active)
        assert np.array_equal(xd.grad.cpu().numpy(), gi_ref)
        _, gw64 = oracle_port.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active)
        assert _gw_close(wd.grad.cpu().numpy(), gw64)


def test_edge_cases(dev, lib, oracle_port, auto_path):
    from torchshifts.functional import shift2d_func
    # empty batch / empty channel set
    for shape in ((0, 3, 4, 4), (2, 0, 4, 4)):
        x = torch.zeros(shape, device=dev, requires_grad=True)
        w = torch.zeros(shape[1], 2, device=dev, requires_grad=True)
        y = shift2d_func(x, w, 0, False)
        assert list(y.shape) == list(shape)
        y.sum().backward()
        assert w.grad.shape == w.shape and float(w.grad.abs().sum()) == 0.0
    # shifts far larger than the tensor, every padding
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 4, 6, 8)).astype(np.float32)
    w = np.array([[1e6, -1e6], [-123456.0, 7.0], [33.5, -0.5], [2.0 ** 40, -(2.0 ** 35)]], dtype=np.float32)
    for pad in range(5):
        for active in (False, True):
            y, _, _ = _run_cuda(dev, 2, x, w, None, pad, active, None)
            assert np.array_equal(y, oracle_port.forward(x, w, pad, active)), (pad, active)
    # borders corner cases resolve to the reference's output shapes
    xb = torch.zeros(1, 1, 16, 16, device=dev)
    wb = torch.zeros(1, 2, device=dev)
    assert list(shift2d_func(xb, wb, 0, False, torch.tensor([[16, 0], [3, 13]])).shape[2:]) == [1, 1]
    with pytest.raises(RuntimeError, match="negative"):
        shift2d_func(xb, wb, 0, False, torch.tensor([[20, 0], [0, 0]]))
    # dtype mismatch and double backward are hard errors, like the reference
    with pytest.raises(RuntimeError, match="scalar type"):
        shift2d_func(xb, wb.double(), 0, False)
    xg = torch.randn(1, 1, 4, 4, device=dev, requires_grad=True)
    wg = torch.randn(1, 2, device=dev, requires_grad=True)
    y = shift2d_func(xg, wg, 0, True)
    (gx,) = torch.autograd.grad(y.sum(), xg, create_graph=True)
    with pytest.raises(RuntimeError, match="double backwards"):
        gx.sum().backward()


def test_modules_and_quantized_convert(dev, oracle_port, auto_path):
    """tests/shifts_test.py of the reference, with assertions instead of prints."""
    import torchshifts
    from torchshifts import Shift2d
    from torchshifts.quantized.modules import Shift2d as QShift2d
    from oracle.oracle import quantize_shift_weights_np
    torch.manual_seed(0)
    channels = 16
    args = {'kernel_size': 3, 'stride': 1, 'padding': (0, 0)}
    ia = torch.rand(8, channels, 64, 64, device=dev, requires_grad=True)
    ib = ia.detach().clone().requires_grad_(True)
    ta = 10 * torch.rand(8, channels, 62, 62, device=dev)
    a = Shift2d(channels, init_shift=1, sparsity_term=0., active_flag=False, emulate_dw=dict(args), init_thumb_rule=2).to(dev)
    b = Shift2d(channels, init_shift=1, sparsity_term=0., active_flag=True, emulate_dw=dict(args), init_thumb_rule=2).to(dev)
    a.weight = b.weight
    oa, la = a(ia)
    ob, lb = b(ib)
    assert la is None and lb is None and oa.shape == ta.shape
    torch.nn.functional.mse_loss(oa, ta).backward()
    ga = a.weight.grad.clone()
    a.weight.grad = None
    torch.nn.functional.mse_loss(ob, ta).backward()
    assert ga is not None and b.weight.grad is not None
    xs, ws = ia.detach().cpu().numpy(), a.weight.detach().cpu().numpy()
    assert np.array_equal(oa.detach().cpu().numpy(), oracle_port.forward(xs, ws, 0, False, [[1, 1], [1, 1]]))
    assert np.array_equal(ob.detach().cpu().numpy(), oracle_port.forward(xs, ws, 0, True, [[1, 1], [1, 1]]))
    # quantized conversion, as torch.quantization.convert would do it
    iq = torch.quantize_per_tensor(ia.detach(), 1 / 255., 0, torch.quint8)
    aq = QShift2d.from_float(a)
    oq = aq(iq)
    raw, wzp = quantize_shift_weights_np(ws)
    want = oracle_port.qforward(iq.int_repr().cpu().numpy(), raw, wzp, 0, 0, [[1, 1], [1, 1]])
    assert np.array_equal(oq.int_repr().cpu().numpy(), want)
    model = torch.nn.Sequential(a)
    model.qconfig = torch.ao.quantization.default_qconfig        # convert only swaps modules that carry a qconfig
    torch.ao.quantization.propagate_qconfig_(model)
    conv = torch.ao.quantization.convert(model, mapping=torchshifts.quant_mapping, inplace=False)
    assert type(conv[0]) is QShift2d
    # sparsity loss and stride-2 emulation (avg-pool reduction)
    c = Shift2d(channels, sparsity_term=5e-4, emulate_dw={'kernel_size': 3, 'stride': 2, 'padding': 1}).to(dev)
    oc, lc = c(ia.detach())
    assert oc.shape[-1] == 32 and torch.allclose(lc, 5e-4 * c.weight.abs().sum())


# ------------------------------------------------------------------------------------------
# BASELINE.json full sizes: properties + oracle on sampled images.
def _sample_check(dev, oracle_port, dim, x, w, g, pad, active, y, gi, n_idx):
    xs = x[n_idx].cpu().numpy()
    ws = w.detach().cpu().numpy()
    assert np.array_equal(y.detach()[n_idx].cpu().numpy(), oracle_port.forward(xs, ws, pad, active))
    if gi is not None:
        gi_ref, _ = oracle_port.backward(g[n_idx].cpu().numpy(), xs, ws, pad, active)
        assert np.array_equal(gi.detach()[n_idx].cpu().numpy(), gi_ref)


def test_full_size_cfg3_shift2d(dev, lib, oracle_port, auto_path):
    """cfg3: N=256 C=256 56x56 fp32, sparse shift, zeros padding, fwd+bwd."""
    from torchshifts.functional import shift2d_func
    torch.manual_seed(0)
    N, C, H, W = 256, 256, 56, 56
    x = torch.randn(N, C, H, W, device=dev)
    g = torch.randn(N, C, H, W, device=dev)
    w = ((torch.rand(C, 2, device=dev) * 2 - 1) * 3).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = shift2d_func(xr, w, 0, False)
    assert lib.ts_last_kernel_path() in BANDWIDTH, "cfg3 must run on a staged (bulk-async / TMA) path"
    y.backward(g)
    assert lib.ts_last_kernel_path() in BANDWIDTH
    gi, gw = xr.grad, w.grad.clone()
    idx = [0, 17, 255]
    _sample_check(dev, oracle_port, 2, x, w, g, 0, False, y, gi, idx)
    # property: zeros-padded integer shift == roll + mask, for the whole tensor
    s = torch.round(w.detach()).long()
    ii = torch.arange(H, device=dev).view(1, H, 1) - s[:, 0].view(C, 1, 1)
    jj = torch.arange(W, device=dev).view(1, 1, W) - s[:, 1].view(C, 1, 1)
    ok = (ii >= 0) & (ii < H) & (jj >= 0) & (jj < W)
    src = (ii.clamp(0, H - 1) * W + jj.clamp(0, W - 1)).view(1, C, H * W).expand(N, C, H * W)
    want = torch.gather(x.view(N, C, H * W), 2, src).view(N, C, H, W) * ok.view(1, C, H, W)
    assert torch.equal(y.detach(), want)
    # property: linearity of the sparse forward (a gather commutes with addition, bit-exactly)
    x2 = torch.randn_like(x)
    assert torch.equal(shift2d_func(x + x2, w.detach(), 0, False), y.detach() + shift2d_func(x2, w.detach(), 0, False))
    del x2, want, src
    # property: grad_weight of the full batch == sum over two half batches (the multi-GPU shard rule)
    gws = []
    for sl in (slice(0, 128), slice(128, 256)):
        wh = w.detach().clone().requires_grad_(True)
        shift2d_func(x[sl], wh, 0, False).backward(g[sl])
        gws.append(wh.grad)
    assert torch.allclose(gws[0] + gws[1], gw, rtol=1e-5, atol=1e-5 * float(gw.abs().max()))
    # grad_weight vs the fp64 oracle on a sub-batch
    wq = w.detach().clone().requires_grad_(True)
    shift2d_func(x[:4], wq, 0, False).backward(g[:4])
    _, gw64 = oracle_port.backward(g[:4].double().cpu().numpy(), x[:4].double().cpu().numpy(), w.detach().double().cpu().numpy(), 0, False)
    assert _gw_close(wq.grad.cpu().numpy(), gw64)
    # determinism: two runs give identical bits (no atomics)
    wd = w.detach().clone().requires_grad_(True)
    xd = x.clone().requires_grad_(True)
    shift2d_func(xd, wd, 0, False).backward(g)
    assert torch.equal(wd.grad, gw) and torch.equal(xd.grad, gi)


def test_full_size_cfg2_shift1d_active_periodic(dev, lib, oracle_port, auto_path):
    """cfg2: N=64 C=512 L=4096, periodic padding, active shift, fp32 and bf16."""
    from torchshifts.functional import shift1d_func
    torch.manual_seed(1)
    N, C, L = 64, 512, 4096
    x = torch.randn(N, C, L, device=dev)
    g = torch.randn(N, C, L, device=dev)
    w = ((torch.rand(C, 1, device=dev) * 2 - 1) * 3).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y = shift1d_func(xr, w, 2, True)
    y.backward(g)
    _sample_check(dev, oracle_port, 1, x, w, g, 2, True, y, xr.grad, [0, 63])
    # property: periodic active shift preserves the per-row sum up to rounding (weights sum to 1)
    assert torch.allclose(y.detach().double().sum(-1), x.double().sum(-1), rtol=0, atol=1e-2)
    # property: periodic sparse shift by s then by -s is the identity
    wi = torch.round(w.detach() * 5)
    assert torch.equal(shift1d_func(shift1d_func(x, wi, 2, False), -wi, 2, False), x)
    wq = w.detach().clone().requires_grad_(True)
    shift1d_func(x[:2], wq, 2, True).backward(g[:2])
    _, gw64 = oracle_port.backward(g[:2].double().cpu().numpy(), x[:2].double().cpu().numpy(), w.detach().double().cpu().numpy(), 2, True)
    assert _gw_close(wq.grad.cpu().numpy(), gw64)
    # bf16 at FULL size (the named configuration): fp32 arithmetic on the bf16 values, one rounding at the store -- so
    # forward and grad_input equal the fp32 oracle rounded once, bit for bit, on the sampled images
    xf, wf, gf = x.bfloat16(), w.detach().bfloat16(), g.bfloat16()
    xfr, wfr = xf.clone().requires_grad_(True), wf.clone().requires_grad_(True)
    yf = shift1d_func(xfr, wfr, 2, True)
    yf.backward(gf)
    for n in (0, 37, 63):
        xs, gs, ws = xf[n:n + 1].float().cpu().numpy(), gf[n:n + 1].float().cpu().numpy(), wf.float().cpu().numpy()
        assert torch.equal(yf[n:n + 1].detach().cpu(), torch.from_numpy(oracle_port.forward(xs, ws, 2, True)).bfloat16()), n
        assert torch.equal(xfr.grad[n:n + 1].cpu(), torch.from_numpy(oracle_port.backward(gs, xs, ws, 2, True)[0]).bfloat16()), n
    del xf, gf, xfr, yf
    # bf16 variant against the fp32 oracle on rounded inputs
    xb, wb, gb = x[:4].bfloat16(), w.detach().bfloat16(), g[:4].bfloat16()
    xbr, wbr = xb.clone().requires_grad_(True), wb.clone().requires_grad_(True)
    yb = shift1d_func(xbr, wbr, 2, True)
    yb.backward(gb)
    y_ref = oracle_port.forward(xb.float().cpu().numpy(), wb.float().cpu().numpy(), 2, True)
    gi_ref, gw_ref = oracle_port.backward(gb.float().cpu().numpy(), xb.float().cpu().numpy(), wb.float().cpu().numpy(), 2, True)
    assert np.allclose(yb.detach().float().cpu().numpy(), y_ref, rtol=1e-2, atol=1e-2)
    assert np.allclose(xbr.grad.float().cpu().numpy(), gi_ref, rtol=1e-2, atol=1e-2)
    assert np.allclose(wbr.grad.float().cpu().numpy(), gw_ref, rtol=1e-2, atol=1e-2 * np.abs(gw_ref).max())


def test_full_size_cfg4_shift3d_active_all_paddings(dev, oracle_port, auto_path):
    """cfg4: N=32 C=128 16x56x56, trilinear active shift, all five paddings."""
    from torchshifts.functional import shift3d_func
    torch.manual_seed(2)
    N, C = 32, 128
    x = torch.randn(N, C, 16, 56, 56, device=dev)
    g = torch.randn(N, C, 16, 56, 56, device=dev)
    w0 = (torch.rand(C, 3, device=dev) * 2 - 1) * 3
    for pad in range(5):
        w = w0.clone().requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        y = shift3d_func(xr, w, pad, True)
        y.backward(g)
        _sample_check(dev, oracle_port, 3, x, w, g, pad, True, y, xr.grad, [pad])
        wq = w0.clone().requires_grad_(True)
        shift3d_func(x[:1], wq, pad, True).backward(g[:1])
        _, gw64 = oracle_port.backward(g[:1].double().cpu().numpy(), x[:1].double().cpu().numpy(), w0.double().cpu().numpy(), pad, True)
        assert _gw_close(wq.grad.cpu().numpy(), gw64), pad
        del y, xr


def test_full_size_cfg5_quantized_shift2d(dev, lib, oracle_port, auto_path):
    """cfg5: qint8 / quint8 Shift2d forward N=256 C=256 56x56, bit-exact."""
    from torchshifts.quantized.functional import shift2d_quantized
    from torchshifts.quantized.modules.shifts import quantize_shift_weights
    from oracle.oracle import quantize_shift_weights_np
    torch.manual_seed(3)
    x = torch.rand(256, 256, 56, 56, device=dev)
    w = (torch.rand(256, 2, device=dev) * 2 - 1) * 3
    raw, wzp = quantize_shift_weights_np(w.cpu().numpy())
    for qdtype, zp in ((torch.qint8, -128), (torch.quint8, 0)):
        xq = torch.quantize_per_tensor(x, 1 / 255., zp, qdtype)
        qw = quantize_shift_weights(w)
        assert np.array_equal(qw.int_repr().cpu().numpy().astype(np.int64), raw)
        yq = shift2d_quantized(xq, qw, 0)
        assert lib.ts_last_kernel_path() in BANDWIDTH
        idx = [0, 100, 255]
        want = oracle_port.qforward(xq.int_repr()[idx].cpu().numpy(), raw, wzp, zp, 0)
        assert np.array_equal(yq.int_repr()[idx].cpu().numpy(), want)
        # checksum property: with zeros padding every output byte is an input byte or the zero point
        hist_in = torch.bincount(xq.int_repr().flatten().to(torch.int64) + 128, minlength=384)
        hist_out = torch.bincount(yq.int_repr().flatten().to(torch.int64) + 128, minlength=384)
        zp_bin = zp + 128
        other = torch.ones_like(hist_in, dtype=torch.bool)
        other[zp_bin] = False
        assert bool((hist_out[other] <= hist_in[other]).all())


def test_cfg1_small_l2_resident(dev, oracle_port, auto_path):
    """cfg1: N=8 C=64 32x32 fp32 SSL zeros -- the reference's own CPU-runnable case, whole tensor vs oracle."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, 64, 32, 32)).astype(np.float32)
    g = rng.standard_normal((8, 64, 32, 32)).astype(np.float32)
    w = ((rng.random((64, 2)) * 2 - 1)).astype(np.float32)
    y, gi, gw = _run_cuda(dev, 2, x, w, g, 0, False, None)
    assert np.array_equal(y, oracle_port.forward(x, w, 0, False))
    gi_ref, _ = oracle_port.backward(g, x, w, 0, False)
    assert np.array_equal(gi, gi_ref)
    _, gw64 = oracle_port.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), 0, False)
    assert _gw_close(gw, gw64)


def test_smoke_in_a_fresh_process():
    """__graft_entry__.smoke() in a new interpreter: the first backward of a process runs on the autograd
    worker thread before that thread has a CUDA context (the TMA path's tensor-map encode is a driver call)."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "[smoke] ok" in out.stdout


def test_more_than_2_31_elements(dev, lib, oracle_port, auto_path):
    """64-bit offsets: a tensor with > 2^31 elements (int8, 2.2 GB); the images at both ends and in the
    middle are checked against the oracle, the rest through a per-image checksum identity."""
    from torchshifts.quantized.functional import shift2d_quantized
    from oracle.oracle import quantize_shift_weights_np
    N, C, H, W = 540, 64, 256, 256
    assert N * C * H * W > 2 ** 31
    gen = torch.Generator(device=dev).manual_seed(11)
    raw = torch.randint(-128, 128, (N, C, H, W), dtype=torch.int8, device=dev, generator=gen)
    xq = torch._make_per_tensor_quantized_tensor(raw, 0.05, 0)            # zero point 0: the TMA family applies
    w = np.linspace(-5, 5, C * 2).reshape(C, 2).astype(np.float32)
    wraw, wzp = quantize_shift_weights_np(w)
    qw = torch._make_per_tensor_quantized_tensor(torch.from_numpy(wraw.astype(np.uint8)).to(dev), 1.0, int(wzp))
    for path in (0, STAGED):
        lib.ts_set_kernel_path(path)
        yq = shift2d_quantized(xq, qw, 0)
        yr = yq.int_repr()
        for n in (0, 269, 539):
            want = oracle_port.qforward(raw[n:n + 1].cpu().numpy(), wraw, wzp, 0, 0)
            assert np.array_equal(yr[n:n + 1].cpu().numpy(), want), (path, n)
        # zeros-padded integer shift: every image loses the same rows / columns, so image n and image 0 agree on
        # (sum of y) - (sum of the surviving window of x); cheap whole-tensor check through per-channel window sums
        sh = wraw.astype(np.int64) - wzp
        c = 7
        sh_h, sh_w = int(sh[c, 0]), int(sh[c, 1])
        win = raw[:, c, max(0, -sh_h):H - max(0, sh_h), max(0, -sh_w):W - max(0, sh_w)]
        assert torch.equal(yr[:, c].to(torch.int32).sum(dim=(1, 2)), win.to(torch.int32).sum(dim=(1, 2)))
        del yq, yr
    lib.ts_set_kernel_path(0)


def test_torch_compile_and_state_dict(dev, auto_path):
    """The ops carry fake kernels and a registered autograd formula: a module with a shift layer traces under
    torch.compile as ONE graph, and its state_dict has the single `weight` parameter."""
    import torchshifts
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Conv2d(8, 8, 1), ).to(dev)
    sh = torchshifts.Shift2d(8, padding='zeros', active_flag=True).to(dev)
    assert list(sh.state_dict().keys()) == ['weight']

    def f(x):
        y, _ = sh(m(x))
        return torch.relu(y)

    x = torch.randn(2, 8, 16, 16, device=dev, requires_grad=True)
    want = f(x)
    want.sum().backward()
    gx, gw = x.grad.clone(), sh.weight.grad.clone()
    x.grad = None; sh.weight.grad = None
    # backend "aot_eager": dynamo + AOTAutograd trace the ops through their fake kernels and the registered
    # autograd formula; no Inductor code generation (this image has no working host compiler for it)
    got = torch.compile(f, dynamic=False, backend="aot_eager", fullgraph=True)(x)
    got.sum().backward()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(x.grad, gx, rtol=1e-5, atol=1e-6) and torch.allclose(sh.weight.grad, gw, rtol=1e-4, atol=1e-5)


def test_torch_compile_cropping_layer_and_inductor(dev, auto_path):
    """A layer built with emulate_dw that CROPS its output (kernel_size > 2*padding+1) and average-pools it
    (stride 2) compiles as one graph: the crop is resolved from Python integers while tracing (the border tensor's
    values cannot be read there).  Backends: aot_eager (always) and inductor (Triton for the pointwise / pooling
    ops around the custom operator)."""
    import torchshifts
    torch.manual_seed(0)
    sh = torchshifts.Shift2d(8, padding='reflect', active_flag=True, emulate_dw={'kernel_size': 3, 'padding': 0, 'stride': 2}).to(dev)
    assert sh.cut_borders is not None

    def f(x):
        y, loss = sh(x * 2.0)
        return torch.relu(y).sum() + loss

    x = torch.randn(2, 8, 16, 16, device=dev, requires_grad=True)
    want = f(x)
    want.backward()
    gx, gw = x.grad.clone(), sh.weight.grad.clone()
    for backend in ("aot_eager", "inductor"):
        x.grad = None; sh.weight.grad = None
        torch._dynamo.reset()
        try:
            got = torch.compile(f, dynamic=False, backend=backend, fullgraph=True)(x)
            got.backward()
        except Exception as e:      # noqa: BLE001
            if backend == "inductor" and "libgomp.spec" in str(e):
                # this image's g++ (the one Inductor picks for host code) cannot link OpenMP: "cannot read spec file
                # 'libgomp.spec'".  An environment defect, not a property of the operators; aot_eager above has passed.
                pytest.skip("Inductor's host compiler is broken in this image (libgomp.spec missing); aot_eager passed")
            raise
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), backend
        assert torch.allclose(x.grad, gx, rtol=1e-5, atol=1e-6), backend
        assert torch.allclose(sh.weight.grad, gw, rtol=1e-4, atol=1e-5), backend


def test_fused_avgpool_epilogue(dev, lib, oracle_port, auto_path):
    """a Shift2d that emulates a stride-2 depth-wise convolution (modules/shifts.py:85-89) shifts, crops and
    average-pools in ONE kernel.  Forward: bit-exact against the oracle's shift followed by a sequential-sum pooling in
    numpy, and against the two-step path (shift op + torch's avg_pool2d); backward: the pooling adjoint + shift backward
    against autograd through the two-step path."""
    import torchshifts
    from torchshifts.functional import shift2d_func
    rng = np.random.default_rng(61)
    # (shape, padding, active, conv kernel, conv padding, fused?)  -- the fused kernel needs input AND output rows of a multiple
    # of 4 elements (a 3x3 "valid" conv crops 1 + 1 columns: never both); such layers take the two-step path inside the
    # same operator and must give the same values (last two cases)
    for shape, pad_name, active, ks, conv_pad, fused in [((2, 8, 16, 16), 'zeros', False, 3, 1, True), ((2, 8, 15, 20), 'reflect', True, 3, 1, True),
                                                         ((3, 4, 18, 28), 'periodic', False, 5, 0, True), ((2, 4, 17, 16), 'symmetric', True, 5, 0, True),
                                                         ((2, 6, 16, 24), 'border', True, 3, 1, True), ((2, 4, 9, 22), 'zeros', True, 3, 1, False),
                                                         ((2, 4, 18, 26), 'reflect', False, 3, 0, False), ((2, 4, 17, 24), 'symmetric', True, 3, 1, True),
                                                         ((2, 4, 19, 44), 'zeros', True, 5, 0, True), ((3, 5, 21, 40), 'periodic', False, 3, 1, True),
                                                         ((2, 3, 56, 56), 'zeros', False, 3, 1, True), ((2, 3, 31, 36), 'reflect', True, 5, 0, True)]:
        torch.manual_seed(3)
        m = torchshifts.Shift2d(shape[1], padding=pad_name, active_flag=active, sparsity_term=0,
                                emulate_dw={'kernel_size': ks, 'stride': 2, 'padding': conv_pad}).to(dev)
        assert (m.cut_borders is not None) == (conv_pad == 0)
        x = rng.standard_normal(shape).astype(np.float32)
        w = m.weight.detach().cpu().numpy()
        pad = torchshifts.modules.shifts.paddings_dict[pad_name]
        borders = m.cut_borders.tolist() if m.cut_borders is not None else None
        xd = torch.from_numpy(x).to(dev).requires_grad_(True)
        out, loss = m(xd)
        assert loss is None and (lib.ts_last_kernel_path() == HALO) == fused, (shape, lib.ts_last_kernel_path())
        y = oracle_port.forward(x, w, pad, active, borders)
        oh, ow = y.shape[2:]
        want = np.zeros(y.shape[:2] + ((oh + 1) // 2, (ow + 1) // 2), np.float32)
        for i in range(want.shape[2]):
            for j in range(want.shape[3]):
                acc, cnt = None, 0
                for di in range(2):
                    for dj in range(2):
                        if 2 * i + di < oh and 2 * j + dj < ow:
                            v = y[:, :, 2 * i + di, 2 * j + dj]
                            acc = v.copy() if acc is None else (acc + v).astype(np.float32)
                            cnt += 1
                want[:, :, i, j] = (acc / np.float32(cnt)).astype(np.float32)
        assert np.array_equal(out.detach().cpu().numpy(), want), (shape, pad_name, active)
        # the two-step path: same values, and its autograd gives the reference gradients
        x2 = torch.from_numpy(x).to(dev).requires_grad_(True)
        w2 = m.weight.detach().clone().requires_grad_(True)
        two = torch.nn.functional.avg_pool2d(shift2d_func(x2, w2, pad, active, m.cut_borders), 2, 2, ceil_mode=True)
        assert torch.equal(two, out)
        g = torch.from_numpy(rng.standard_normal(want.shape).astype(np.float32)).to(dev)
        before = lib.ts_launch_count()
        out.backward(g)
        # the fused backward (pooling adjoint applied while the gradient is staged: needs pooled rows of a multiple of 16 bytes)
        # is ONE shift kernel + the pass-2 reduction; otherwise ATen's avg_pool2d_backward runs first
        if fused and ow % 8 == 0:
            assert lib.ts_launch_count() - before == 2, (shape, lib.ts_launch_count() - before)
            # (the kernel path is recorded per thread and autograd ran the backward on its own: ask the operator directly)
            std = torch.tensor([0, shape[2], 0, shape[3], 0, 1], dtype=torch.int32)
            if borders is not None:
                std = torch.tensor([borders[0][0], shape[2] - borders[0][1], borders[1][0], shape[3] - borders[1][1], 0, 1], dtype=torch.int32)
            lib.ts_set_kernel_path(1); lib.ts_set_kernel_path(0)
            gi_d, gw_d = torch.ops.torchshifts._shift2d_avgpool2_backward(g, m.weight.detach(), xd.detach(), std, list(y.shape), pad, active)
            assert lib.ts_last_kernel_path() == HALO, (shape, lib.ts_last_kernel_path())
            assert torch.equal(gi_d, xd.grad) and torch.equal(gw_d, m.weight.grad)
        two.backward(g)
        assert torch.equal(xd.grad, x2.grad), (shape, pad_name, active)
        assert torch.allclose(m.weight.grad, w2.grad, rtol=1e-5, atol=1e-6)
        # and against the oracle: pooling adjoint in numpy (grad / count to every element of the window), then the shift backward
        gy = np.zeros(y.shape, np.float32)
        gp = g.cpu().numpy()
        for i in range(oh):
            cnt = np.float32((2 if 2 * (i // 2) + 1 < oh else 1) * 2)
            gy[:, :, i, :] = np.repeat(gp[:, :, i // 2, :], 2, axis=-1)[..., :ow] / cnt
        if ow % 2 == 0:
            gi_ref, _ = oracle_port.backward(gy, x, w, pad, active, borders)
            assert np.array_equal(xd.grad.cpu().numpy(), gi_ref), (shape, pad_name, active)


def test_cuda_graph_capture_and_replay(dev, lib, oracle_port, auto_path):
    """The C ABI only enqueues work on the caller's stream (no sync, no allocation, tensor maps are
    encoded on the host and passed by value), so forward + backward capture into a CUDA graph; the
    replay on new data matches the oracle.  cfg1-sized layers are launch-bound without it."""
    fwd = torch.ops.torchshifts._shift2d_forward
    bwd = torch.ops.torchshifts._shift2d_backward
    rng = np.random.default_rng(4)
    shape = (8, 64, 32, 32)
    borders = torch.tensor([0, 32, 0, 32, 0, 1], dtype=torch.int32)
    x = torch.zeros(shape, device=dev)
    g = torch.zeros(shape, device=dev)
    w = torch.from_numpy(((rng.random((64, 2)) * 2 - 1) * 2).astype(np.float32)).to(dev)
    for pad, active in ((0, False), (3, True)):
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                                   # warm-up outside the capture
                y = fwd(x, w, borders, list(shape), pad, active)
                gi, gw = bwd(g, w, x, borders, pad, active)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            y = fwd(x, w, borders, list(shape), pad, active)
            gi, gw = bwd(g, w, x, borders, pad, active)
        for seed in (1, 2):
            xs = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
            gs = np.random.default_rng(seed + 10).standard_normal(shape).astype(np.float32)
            x.copy_(torch.from_numpy(xs)); g.copy_(torch.from_numpy(gs))
            graph.replay()
            torch.cuda.synchronize()
            wn = w.cpu().numpy()
            assert np.array_equal(y.cpu().numpy(), oracle_port.forward(xs, wn, pad, active))
            gi_ref, _ = oracle_port.backward(gs, xs, wn, pad, active)
            assert np.array_equal(gi.cpu().numpy(), gi_ref)
            _, gw64 = oracle_port.backward(gs.astype(np.float64), xs.astype(np.float64), wn.astype(np.float64), pad, active)
            assert _gw_close(gw.cpu().numpy(), gw64)


def test_quantized_channels_last_keeps_its_memory_format(dev, oracle_port, auto_path):
    """quantized/shifts_quantized.cpp:119-122: the output takes the input's memory format."""
    from torchshifts.quantized.functional import shift2d_quantized
    from oracle.oracle import quantize_shift_weights_np
    rng = np.random.default_rng(31)
    x = torch.from_numpy(rng.random((2, 8, 6, 16)).astype(np.float32)).to(dev)
    w = ((rng.random((8, 2)) * 2 - 1) * 2).astype(np.float32)
    wraw, wzp = quantize_shift_weights_np(w)
    qw = torch._make_per_tensor_quantized_tensor(torch.from_numpy(wraw.astype(np.uint8)).to(dev), 1.0, int(wzp))
    for qdtype, zp in ((torch.quint8, 3), (torch.qint8, -5)):
        xq = torch.quantize_per_tensor(x, 0.01, zp, qdtype)
        xcl = xq.contiguous(memory_format=torch.channels_last)
        assert xcl.is_contiguous(memory_format=torch.channels_last) and not xcl.is_contiguous()
        for pad in (0, 3):
            y = shift2d_quantized(xcl, qw, pad)
            assert y.is_contiguous(memory_format=torch.channels_last) and y.dtype == qdtype
            assert y.q_scale() == xq.q_scale() and y.q_zero_point() == zp
            want = oracle_port.qforward(xq.int_repr().cpu().numpy(), wraw, wzp, zp, pad)
            assert np.array_equal(y.int_repr().cpu().numpy(), want)
            assert shift2d_quantized(xq, qw, pad).is_contiguous()


def test_misaligned_dense_tensors_fall_back(dev, lib, oracle_port, auto_path):
    """Dense NCHW tensors whose base pointer is not 16-byte aligned (a view into a larger buffer) cannot use
    bulk / TMA copies: the planners must reject them and the generic family must give the same values."""
    from torchshifts.functional import shift2d_func
    rng = np.random.default_rng(41)
    shape = (2, 3, 8, 16)
    n = int(np.prod(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    w = ((rng.random((3, 2)) * 2 - 1) * 2).astype(np.float32)
    for pad, active in ((0, False), (0, True), (2, True)):
        buf = torch.zeros(n + 1, device=dev)
        xd = buf[1:].view(shape)
        xd.copy_(torch.from_numpy(x))
        assert xd.data_ptr() % 16 != 0 and xd.is_contiguous()
        xd.requires_grad_(True)
        wd = torch.from_numpy(w).to(dev).requires_grad_(True)
        gbuf = torch.zeros(n + 3, device=dev)
        gd = gbuf[3:].view(shape)
        gd.copy_(torch.from_numpy(g))
        y = shift2d_func(xd, wd, pad, active)
        assert lib.ts_last_kernel_path() == GENERIC
        y.backward(gd)
        assert np.array_equal(y.detach().cpu().numpy(), oracle_port.forward(x, w, pad, active))
        gi_ref, _ = oracle_port.backward(g, x, w, pad, active)
        assert np.array_equal(xd.grad.cpu().numpy(), gi_ref)
        _, gw64 = oracle_port.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active)
        assert _gw_close(wd.grad.cpu().numpy(), gw64)


def test_fused_allreduce_single_rank(dev, lib, oracle_port, auto_path):
    """ts_shift_backward_allreduce with a one-rank peer group: the exchange protocol ({epoch:value} words stored
    into the own buffer, poll, rank-ordered sum) must reproduce the plain backward; several calls in a row
    exercise the device-side call counter / double buffering.  (Multi-rank: test_fused_allreduce_two_ranks.)"""
    from torchshifts.functional import shift2d_func, shift3d_func
    from torchshifts.sharded import FusedGradWeightAllReduce
    rng = np.random.default_rng(51)
    fused = FusedGradWeightAllReduce(capacity=512, device=dev)
    for it, (fn, shape, pad, active) in enumerate([(shift2d_func, (4, 6, 16, 16), 0, False), (shift2d_func, (4, 6, 16, 16), 3, True),
                                                   (shift3d_func, (2, 3, 4, 8, 8), 2, True), (shift2d_func, (3, 5, 7, 9), 4, False),
                                                   (shift2d_func, (4, 6, 16, 16), 0, True)]):
        dim = len(shape) - 2
        x = rng.standard_normal(shape).astype(np.float32)
        g = rng.standard_normal(shape).astype(np.float32)
        w = ((rng.random((shape[1], dim)) * 2 - 1) * 2).astype(np.float32)
        xd = torch.from_numpy(x).to(dev).requires_grad_(True)
        wd = torch.from_numpy(w).to(dev).requires_grad_(True)
        with fused:
            fn(xd, wd, pad, active).backward(torch.from_numpy(g).to(dev))
        assert fused.calls() == it + 1
        gi_ref, _ = oracle_port.backward(g, x, w, pad, active)
        _, gw64 = oracle_port.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active)
        assert np.array_equal(xd.grad.cpu().numpy(), gi_ref)
        assert _gw_close(wd.grad.cpu().numpy(), gw64), (it, wd.grad.cpu().numpy(), gw64)
    # outside the context the plain path runs again
    xd = torch.from_numpy(x).to(dev).requires_grad_(True)
    wd = torch.from_numpy(w).to(dev).requires_grad_(True)
    shift2d_func(xd, wd, 0, True).backward(torch.from_numpy(g).to(dev))
    assert fused.calls() == 5 and _gw_close(wd.grad.cpu().numpy(), gw64)


def test_fused_allreduce_two_ranks():
    """The in-kernel exchange on TWO ranks (needs >= 2 visible GPUs, skipped otherwise): tools/fused_allreduce_probe.py
    under torchrun -- eager mixed layers, a ragged batch with an empty shard, CUDA-graph replay; every result against
    NCCL's all-reduce of the plain backward (rtol 1e-5) and bit-identical across the ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run on the box with: gpurun --gpus 2 -- python -m pytest tests -m gpu -k two_ranks)")
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tools" / "fused_allreduce_probe.py"), "--quick"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PROBE OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


NHWC = 4


def test_quantized_channels_last_native_kernel(dev, lib, oracle_port, auto_path):
    """channels-last quantized tensors go through ts_qshift_forward_nhwc (NHWC in, NHWC out,
    no layout conversion), bit-exact with the oracle for every padding, dtype, crop and channel count
    (kernels/shifts_kernels.h:574-624)."""
    from torchshifts.quantized.functional import shift2d_quantized, shift3d_quantized
    rng = np.random.default_rng(41)
    cases = [((2, 8, 6, 16), None), ((3, 3, 5, 7), None), ((2, 20, 9, 12), [[1, 0], [2, 1]]), ((1, 1028, 4, 6), None),
             ((2, 64, 56, 56), None), ((2, 6, 3, 4, 5), None), ((1, 16, 6, 10, 12), [[0, 1], [1, 0], [2, 1]]),
             ((4, 256, 1, 40), None), ((2, 128, 20, 24), [[2, 1], [0, 3]]), ((1, 96, 33, 17), None), ((3, 32, 8, 8), None),
             ((300, 128, 9, 5), None), ((3, 256, 30, 14), None), ((2, 512, 12, 8), [[1, 2], [0, 0]]), ((5, 128, 40, 12), [[3, 0], [1, 1]]), ((2, 384, 16, 10), None), ((1, 1024, 12, 6), [[0, 1], [1, 0]])]
    for shape, borders in cases:
        dim = len(shape) - 2
        fn = shift2d_quantized if dim == 2 else shift3d_quantized
        fmt = torch.channels_last if dim == 2 else torch.channels_last_3d
        wraw = (rng.integers(-5, 5, size=(shape[1], dim), endpoint=True) + 128).astype(np.uint8)
        qw = torch._make_per_tensor_quantized_tensor(torch.from_numpy(wraw).to(dev), 1.0, 128)
        cut = None if borders is None else torch.tensor(borders)
        for qdtype, npdt, zp in ((torch.quint8, np.uint8, 3), (torch.qint8, np.int8, -5), (torch.qint32, np.int32, 70000)):
            info = np.iinfo(npdt)
            raw = rng.integers(max(info.min, -2 ** 20), min(info.max, 2 ** 20), size=shape, endpoint=True).astype(npdt)
            xq = torch._make_per_tensor_quantized_tensor(torch.from_numpy(raw).to(dev), 0.02, zp)
            to_cl = (0,) + tuple(range(2, dim + 2)) + (1,)
            to_nc = (0, dim + 1) + tuple(range(1, dim + 1))
            raw_cl = np.ascontiguousarray(raw.transpose(to_cl))
            xcl = torch._make_per_tensor_quantized_tensor(torch.from_numpy(raw_cl).to(dev), 0.02, zp).permute(*to_nc)
            assert xcl.is_contiguous(memory_format=fmt) and tuple(xcl.shape) == shape
            if xcl.is_contiguous():
                continue                          # degenerate shapes where both layouts coincide
            # kernel variants: automatic choice, the direct (L1) kernel, and -- where it applies (2-D, 1-byte
            # elements, C % 32 == 0) -- the shared-memory ring kernel, also with a ring too small for the shifts
            variants = [b"nhwc_variant=0,nhwc_ring_rows=0", b"nhwc_variant=1,nhwc_ring_rows=0"]
            if dim == 2 and npdt is not np.int32 and shape[1] % 32 == 0:
                variants += [b"nhwc_variant=2,nhwc_ring_rows=0", b"nhwc_variant=2,nhwc_ring_rows=3"]
            # the row-pipelined kernel (dense pixels, C = 128 / 256 / 512), also with a ring of 9 rows: windows of 2 rows,
            # so most taps of the +-5 shifts are outside their window and come from global memory
            if dim == 2 and npdt is not np.int32 and shape[1] % 128 == 0:
                variants += [b"nhwc_variant=3,nhwc_ring_rows=0", b"nhwc_variant=3,nhwc_ring_rows=9"]
            for pad in range(5):
                want = oracle_port.qforward(raw, wraw.astype(np.int64), 128, zp, pad, borders)
                for variant in variants:
                    assert lib.ts_set_tuning(variant) == 0
                    y = fn(xcl, qw, pad, cut)
                    assert lib.ts_last_kernel_path() == NHWC, (shape, pad, variant)
                    assert y.is_contiguous(memory_format=fmt) and y.dtype == qdtype
                    assert y.q_scale() == xq.q_scale() and y.q_zero_point() == zp
                    got = y.permute(*to_cl).int_repr().cpu().numpy().transpose(to_nc)
                    assert np.array_equal(got, want), (shape, borders, qdtype, pad, variant)
                lib.ts_set_tuning(b"nhwc_variant=0,nhwc_ring_rows=0")
                # the planar path on the same tensor gives the same integers
                assert np.array_equal(fn(xq, qw, pad, cut).int_repr().cpu().numpy(), want)


def test_quantized_degenerate_zero_scale_weights(dev, oracle_port, auto_path):
    """all-equal float weights give quantize_shift_weights a scale of 0.  torch then stores an integer
    representation of all 0 when the weights are quantized on the CPU (where the reference's quantized modules live: every
    channel shifts by 0 - 128 = -128) and all 255 when they are quantized on a CUDA device (+127); the kernels consume
    whatever integers the tensor holds, exactly like the reference's `weights.int_repr().long() - zero_point`."""
    from torchshifts.quantized.functional import shift2d_quantized
    from torchshifts.quantized.modules.shifts import quantize_shift_weights
    from oracle.oracle import quantize_shift_weights_np
    rng = np.random.default_rng(77)
    w = torch.full((6, 2), 1.25)
    raw_cpu, wzp = quantize_shift_weights_np(w.numpy())
    assert wzp == 128 and not raw_cpu.any()
    xr = rng.integers(0, 255, size=(2, 6, 12, 20), endpoint=True).astype(np.uint8)
    xq = torch._make_per_tensor_quantized_tensor(torch.from_numpy(xr).to(dev), 0.05, 7)
    for qw in (quantize_shift_weights(w).to(dev), quantize_shift_weights(w.to(dev))):
        assert qw.q_scale() == 0 and qw.q_zero_point() == 128
        raw = qw.int_repr().cpu().numpy().astype(np.int64)
        assert len(np.unique(raw)) == 1 and int(raw.flat[0]) in (0, 255)
        for pad in range(5):
            want = oracle_port.qforward(xr, raw, 128, 7, pad)
            for t in (xq, xq.contiguous(memory_format=torch.channels_last)):
                assert np.array_equal(shift2d_quantized(t, qw, pad).int_repr().cpu().numpy(), want), (pad, int(raw.flat[0]))
    assert not quantize_shift_weights(w).int_repr().any()          # the CPU behaviour SURVEY a15 records


def test_full_size_cfg5_channels_last(dev, lib, oracle_port, auto_path):
    """cfg5 in channels-last: N=256 C=256 56x56 qint8, native NHWC kernel == planar kernel on every byte,
    == oracle on sampled images."""
    from torchshifts.quantized.functional import shift2d_quantized
    from torchshifts.quantized.modules.shifts import quantize_shift_weights
    from oracle.oracle import quantize_shift_weights_np
    torch.manual_seed(5)
    x = torch.rand(256, 256, 56, 56, device=dev)
    w = (torch.rand(256, 2, device=dev) * 2 - 1) * 3
    raw, wzp = quantize_shift_weights_np(w.cpu().numpy())
    qw = quantize_shift_weights(w)
    xq = torch.quantize_per_tensor(x, 1 / 255., -128, torch.qint8)
    del x
    xcl = xq.contiguous(memory_format=torch.channels_last)
    for pad in (0, 4):
        planar = shift2d_quantized(xq, qw, pad).int_repr()
        idx = [0, 131, 255]
        want = oracle_port.qforward(xq.int_repr()[idx].cpu().numpy(), raw, wzp, -128, pad)
        for variant in (b"nhwc_variant=3", b"nhwc_variant=2", b"nhwc_variant=1", b"nhwc_variant=0"):       # row-pipelined, ring, direct, automatic
            assert lib.ts_set_tuning(variant) == 0
            y = shift2d_quantized(xcl, qw, pad)
            assert lib.ts_last_kernel_path() == NHWC and y.is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(y.int_repr(), planar), (pad, variant)
            assert np.array_equal(y.int_repr()[idx].cpu().numpy(), want)


def test_channels_last_to_planar_adapter(dev, lib, oracle_port, auto_path):
    """ts_nhwc_to_nchw (the float path's layout pass for channels-last inputs) == torch's .contiguous(), bit for bit,
    for ragged tile edges and every element size; and a channels-last float input gives the oracle's result through it."""
    import ctypes as ct
    from torchshifts.functional import shift2d_func
    torch.manual_seed(11)
    stream = ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for dtype in (torch.float32, torch.bfloat16, torch.float16, torch.float64):
        for shape in ((2, 5, 7, 9), (3, 64, 56, 56), (1, 33, 3, 4, 5), (2, 256, 1, 40), (1, 1, 6, 6), (2, 70, 31, 1)):
            fmt = torch.channels_last if len(shape) == 4 else torch.channels_last_3d
            x = torch.randn(shape, device=dev).to(dtype).contiguous(memory_format=fmt)
            out = torch.full(shape, 7, dtype=dtype, device=dev)
            n, c = shape[0], shape[1]
            assert lib.ts_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), n, c, x.numel() // (n * c), x.element_size(), stream) == 0
            assert out.is_contiguous() and torch.equal(out, x.contiguous()), (dtype, shape)
    assert lib.ts_nhwc_to_nchw(0, 0, 1, 1, 1, 3, stream) != 0            # null pointers / odd element size are refused
    x = torch.randn(4, 96, 40, 40, device=dev)                              # above the copy threshold: _dense() uses the adapter
    w = (torch.rand(96, 2, device=dev) * 2 - 1) * 2
    before = lib.ts_launch_count()
    y = shift2d_func(x.contiguous(memory_format=torch.channels_last), w, 3, True)
    assert lib.ts_launch_count() - before == 2 and lib.ts_last_kernel_path() in BANDWIDTH
    assert y.is_contiguous()
    assert np.array_equal(y.cpu().numpy(), oracle_port.forward(x.cpu().numpy(), w.cpu().numpy(), 3, True))
