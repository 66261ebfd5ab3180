"""ctypes/numpy front-end of the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``)
may import this module.  The product package (``activesparseshifts-pytorch_b200/torchshifts``)
never does: it has no CPU compute path at all.

Two libraries sit behind the same numpy API:

* ``Oracle("port")``      -> ``oracle/liboracle_shifts.so`` -- our plain-C restatement
  (``shift_oracle.c``; cites the reference lines it follows).
* ``Oracle("reference")`` -> ``oracle/_ref/libref_shifts.so`` -- the reference's own per-element
  headers (``kernels/shifts_kernels.h``) compiled from /root/reference behind a C ABI
  (``ref_driver.cpp``).  The weight split for this one is done here in numpy with the same
  operations as ``cpu/shifts_cpu.cpp:223-224`` (forward) and ``:242-244`` (backward).

Arrays are ``[N, C, S0(, S1(, S2))]``; weights ``[C, dim]``; borders ``[dim, 2]`` of
``(left_cut, right_cut)`` as in ``torchshifts/functional.py`` or ``None``.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "liboracle_shifts.so"
REF_LIB = HERE / "_ref" / "libref_shifts.so"
REF_FULL_LIB = HERE / "_ref" / "torchshifts_ref" / "_C.so"

_i64p = ct.POINTER(ct.c_int64)


def build(full: bool = False) -> None:
    """Compile the checkers (gcc / g++; seconds).  ``ref`` is skipped when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=False)
    if full and not REF_FULL_LIB.exists() and Path("/root/reference/setup.py").exists():
        subprocess.run([str(HERE / "build_ref_full.sh")], check=False)


def _arr_i64(vals) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(vals, dtype=np.int64))


def _p(a: np.ndarray):
    return a.ctypes.data_as(ct.c_void_p)


def _geometry(x: np.ndarray, dim: int):
    assert x.ndim == dim + 2, f"expected a {dim + 2}-D array, got shape {x.shape}"
    S = list(x.shape[2:]) + [1] * (3 - dim)
    isz = x.itemsize
    xs = [s // isz for s in x.strides] + [0] * (3 - dim)
    return _arr_i64(S), _arr_i64(xs)


def check_borders(dim: int, sizes, user=None):
    """Restatement of csrc/ops/shifts.cpp:93-135 -> (lb[3], rb[3]) in input coordinates."""
    S = list(sizes) + [1] * (3 - dim)
    lb = [0, 0, 0]
    rb = [S[a] if a < dim else 1 for a in range(3)]
    if user is not None and np.asarray(user).size != 0:
        u = np.asarray(user).astype(np.int32).reshape(-1, 2)
        for a in range(dim):
            size = int(S[a])
            r = rb[a] - int(u[a, 1])
            l = int(u[a, 0])
            if r - l < 1:
                r = l + 1
            if l == size:
                l = size - 1
                r = l + 1
            if r == 0:
                l, r = 0, 1
            l = max(0, l)
            r = min(size, r)
            if r - l < 0:
                raise RuntimeError("check_borders: negative output dimension (the reference crashes here too)")
            lb[a], rb[a] = l, r
    return lb, rb


def split_forward(w: np.ndarray, active: bool):
    """cpu/shifts_cpu.cpp:223-224."""
    iw = (np.floor(w) if active else np.rint(w)).astype(np.int64)
    dw = (w - iw.astype(w.dtype)) if active else np.zeros_like(w)
    return np.ascontiguousarray(iw), np.ascontiguousarray(dw.astype(w.dtype))


def split_backward(w: np.ndarray, active: bool):
    """cpu/shifts_cpu.cpp:242-244."""
    if active:
        dw = w - np.floor(w)
        iw = (w - dw).astype(np.int64)          # truncating cast, like .to(kLong)
    else:
        dw = np.where(w > 0, w - np.floor(w), np.ceil(w) - w)
        iw = np.rint(w).astype(np.int64)
    return np.ascontiguousarray(iw), np.ascontiguousarray(dw.astype(w.dtype))


class Oracle:
    def __init__(self, kind: str = "port", threads: int = 1):
        assert kind in ("port", "reference")
        self.kind = kind
        self.threads = int(threads)
        path = PORT_LIB if kind == "port" else REF_LIB
        if not path.exists():
            build()
        if not path.exists():
            raise FileNotFoundError(f"{path} is missing (run `make -C oracle`)")
        self.lib = ct.CDLL(str(path))
        if kind == "reference":
            self.lib.ref_max_threads.restype = ct.c_int

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def available(kind: str) -> bool:
        return (PORT_LIB if kind == "port" else REF_LIB).exists()

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads()) if self.kind == "reference" else 1

    @staticmethod
    def _sfx(dtype) -> str:
        if dtype == np.float32:
            return "f32"
        if dtype == np.float64:
            return "f64"
        raise TypeError(f"oracle supports float32/float64 (the reference CPU dispatch, shifts_cpu.cpp:228), got {dtype}")

    # ------------------------------------------------------------------ float forward
    def forward(self, x: np.ndarray, w: np.ndarray, pad: int, active: bool, borders=None) -> np.ndarray:
        dim = x.ndim - 2
        sfx = self._sfx(x.dtype)
        assert w.dtype == x.dtype and w.shape == (x.shape[1], dim)
        w = np.ascontiguousarray(w)
        S, xs = _geometry(x, dim)
        lb, rb = check_borders(dim, x.shape[2:], borders)
        out_shape = list(x.shape[:2]) + [rb[a] - lb[a] for a in range(dim)]
        y = np.empty(out_shape, dtype=x.dtype)
        N, C = x.shape[:2]
        lb_a, rb_a = _arr_i64(lb), _arr_i64(rb)
        if self.kind == "port":
            fn = getattr(self.lib, f"oracle_shift_forward_{sfx}")
            fn.restype = None
            fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(int(active)), _p(x), _p(xs), _p(y), _p(w),
               ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a))
        else:
            iw, dw = split_forward(w, active)
            fn = getattr(self.lib, f"ref_shift_forward_{sfx}")
            fn.restype = ct.c_int
            rc = fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(int(active)), _p(x), _p(xs), _p(y), _p(iw), _p(dw),
                    ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a), ct.c_int(self.threads))
            assert rc == 0
        return y

    # ------------------------------------------------------------------ float backward
    def backward(self, grad: np.ndarray, x: np.ndarray, w: np.ndarray, pad: int, active: bool, borders=None):
        dim = x.ndim - 2
        sfx = self._sfx(x.dtype)
        assert grad.dtype == x.dtype and w.dtype == x.dtype
        w = np.ascontiguousarray(w)
        grad = np.ascontiguousarray(grad)
        S, xs = _geometry(x, dim)
        lb, rb = check_borders(dim, x.shape[2:], borders)
        assert list(grad.shape[2:]) == [rb[a] - lb[a] for a in range(dim)], "grad shape does not match the cropped output"
        gi = np.empty(x.shape, dtype=x.dtype)
        gw = np.zeros(w.shape, dtype=x.dtype)
        N, C = x.shape[:2]
        lb_a, rb_a = _arr_i64(lb), _arr_i64(rb)
        if self.kind == "port":
            fn = getattr(self.lib, f"oracle_shift_backward_{sfx}")
            fn.restype = None
            fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(int(active)), _p(grad), _p(x), _p(xs), _p(gi), _p(gw), _p(w),
               ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a))
        else:
            iw, dw = split_backward(w, active)
            fn = getattr(self.lib, f"ref_shift_backward_{sfx}")
            fn.restype = ct.c_int
            rc = fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(int(active)), _p(grad), _p(x), _p(xs), _p(gi), _p(gw),
                    _p(iw), _p(dw), ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a), ct.c_int(self.threads))
            assert rc == 0
        return gi, gw

    # ------------------------------------------------------------------ quantized forward
    def qforward(self, xraw: np.ndarray, wq: np.ndarray, wzp: int, zp: int, pad: int, borders=None) -> np.ndarray:
        """xraw: raw integer representation (int8/uint8/int32); wq: raw integer weights [C, dim]."""
        dim = xraw.ndim - 2
        esize = xraw.itemsize
        assert esize in (1, 4)
        wq = _arr_i64(wq)
        assert wq.shape == (xraw.shape[1], dim)
        S, xs = _geometry(xraw, dim)
        lb, rb = check_borders(dim, xraw.shape[2:], borders)
        out_shape = list(xraw.shape[:2]) + [rb[a] - lb[a] for a in range(dim)]
        y = np.empty(out_shape, dtype=xraw.dtype)
        N, C = xraw.shape[:2]
        lb_a, rb_a = _arr_i64(lb), _arr_i64(rb)
        if self.kind == "port":
            fn = self.lib.oracle_qshift_forward
            fn.restype = None
            fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(esize), _p(xraw), _p(xs), _p(y), _p(wq), ct.c_int64(wzp),
               ct.c_int64(zp), ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a))
        else:
            fn = self.lib.ref_qshift_forward
            fn.restype = ct.c_int
            rc = fn(ct.c_int(dim), ct.c_int(pad), ct.c_int(esize), _p(xraw), _p(xs), _p(y), _p(wq), ct.c_int64(wzp),
                    ct.c_int64(zp), ct.c_int64(N), ct.c_int64(C), _p(S), _p(lb_a), _p(rb_a), ct.c_int(self.threads))
            assert rc == 0
        return y

    def remap_axis(self, pad: int, length: int, idx) -> np.ndarray:
        assert self.kind == "port"
        idx = _arr_i64(idx)
        out = np.empty_like(idx)
        self.lib.oracle_remap_axis.restype = None
        self.lib.oracle_remap_axis(ct.c_int(pad), ct.c_int64(length), ct.c_int64(idx.size), _p(idx), _p(out))
        return out


def quantize_shift_weights_np(w: np.ndarray):
    """Integer shifts of torchshifts/quantized/modules/shifts.py:10-12 without torch:
    scale = ceil((max-min)/255), zero point 128, quint8 -> (raw uint8 [C,dim], zero_point)."""
    import math
    scale = math.ceil((float(w.max()) - float(w.min())) / 255.0)
    if scale == 0:
        raw = np.zeros(w.shape, dtype=np.int64)  # degenerate reference behaviour, SURVEY.md a15
    else:
        raw = np.clip(np.rint(w.astype(np.float32) / np.float32(scale)) + 128, 0, 255).astype(np.int64)
    return raw, 128
