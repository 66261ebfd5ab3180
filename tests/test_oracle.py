"""Pin the CPU oracle (oracle/shift_oracle.c) before anything is checked against it.

Sources of truth, in order of authority:
  1. tests/golden/*.npz -- outputs of the reference's own torch extension (make_golden.py);
  2. the known-answer table of SURVEY.md 8(c), typed in here by hand;
  3. oracle/_ref/libref_shifts.so -- the reference's per-element headers behind a C ABI.
All comparisons are bit-exact (np.array_equal): same arithmetic, same order, no FMA.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle, check_borders, quantize_shift_weights_np, split_backward, split_forward

BORDERS = {1: [[1, 2]], 2: [[1, 1], [2, 1]], 3: [[1, 0], [0, 2], [1, 1]]}

# shift1d_func([10,11,12,13,14], [[s]], pad, False); columns zeros..symmetric
KAT = {
    -7: ["0 0 0 0 0", "14 14 14 14 14", "12 13 14 10 11", "11 10 11 12 13", "12 11 10 10 11"],
    -4: ["14 0 0 0 0", "14 14 14 14 14", "14 10 11 12 13", "14 13 12 11 10", "14 14 13 12 11"],
    -1: ["11 12 13 14 0", "11 12 13 14 14", "11 12 13 14 10", "11 12 13 14 13", "11 12 13 14 14"],
    0: ["10 11 12 13 14"] * 5,
    1: ["0 10 11 12 13", "10 10 11 12 13", "14 10 11 12 13", "11 10 11 12 13", "10 10 11 12 13"],
    2: ["0 0 10 11 12", "10 10 10 11 12", "13 14 10 11 12", "12 11 10 11 12", "11 10 10 11 12"],
    4: ["0 0 0 0 10", "10 10 10 10 10", "11 12 13 14 10", "14 13 12 11 10", "13 12 11 10 10"],
    5: ["0 0 0 0 0", "10 10 10 10 10", "10 11 12 13 14", "13 14 13 12 11", "14 13 12 11 10"],
    7: ["0 0 0 0 0", "10 10 10 10 10", "13 14 10 11 12", "11 12 13 14 13", "13 14 14 13 12"],
}


def _case_args(name):
    parts = name.split("_")
    if parts[0] == "cl":
        dim, pad, active = int(parts[1][1]), int(parts[2][1]), int(parts[3][1])
        return dim, pad, bool(active), None
    dim, use_b, pad, active = int(parts[1][1]), int(parts[3][1]), int(parts[4][1]), int(parts[5][1])
    return dim, pad, bool(active), (BORDERS[dim] if use_b else None)


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_known_answer_table(kind):
    if not Oracle.available(kind):
        pytest.skip(f"{kind} library not built")
    orc = Oracle(kind)
    x = np.array([10, 11, 12, 13, 14], dtype=np.float32).reshape(1, 1, 5)
    for s, rows in KAT.items():
        for pad, row in enumerate(rows):
            y = orc.forward(x, np.array([[s]], dtype=np.float32), pad, False)
            assert y.reshape(-1).astype(int).tolist() == [int(t) for t in row.split()], (s, pad)


def test_kat_fixture_agrees_with_survey_table(golden):
    kat = golden["kat"]
    for si, s in enumerate(kat["shift1d_shifts"].tolist()):
        for pad in range(5):
            assert kat["shift1d_table"][si, pad].astype(int).tolist() == [int(t) for t in KAT[s][pad].split()]


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_more_known_answers(kind, golden):
    if not Oracle.available(kind):
        pytest.skip(f"{kind} library not built")
    orc, kat = Oracle(kind), golden["kat"]
    x2 = np.arange(120, dtype=np.float32).reshape(2, 3, 4, 5)
    w2 = np.array([[1., 0.], [0., -1.], [0.4, 1.6]], dtype=np.float32)
    y = orc.forward(x2, w2, 0, False)
    assert np.array_equal(y, kat["shift2d_arange_y"])
    assert np.array_equal(y[0, 0, 0], np.zeros(5)) and np.array_equal(y[0, 0, 1:], x2[0, 0, :-1])   # rows down by 1
    assert np.array_equal(y[0, 1, :, :-1], x2[0, 1, :, 1:]) and not y[0, 1, :, -1].any()            # cols left by 1
    assert np.array_equal(y[0, 2, :, 2:], x2[0, 2, :, :-2])                                           # round(1.6) = 2
    # half-to-even ties (torch::round on the CPU path, cpu/shifts_cpu.cpp:223)
    xt = np.arange(42, dtype=np.float32).reshape(1, 7, 6)
    assert np.array_equal(orc.forward(xt, kat["ties_w"].astype(np.float32), 2, False), kat["ties_y"])
    # a size-1 axis ignores its shift (kernels/shifts_kernels.h:40)
    y = orc.forward(np.arange(4, dtype=np.float32).reshape(1, 1, 1, 4), np.array([[3., 1.]], dtype=np.float32), 3, False)
    assert np.array_equal(y, kat["size1_reflect_y"]) and y.reshape(-1).tolist() == [1, 0, 1, 2]
    y = orc.forward(np.arange(8, dtype=np.float32).reshape(1, 1, 2, 4), np.array([[1., 0.]], dtype=np.float32), 3, False)
    assert np.array_equal(y, kat["len2_reflect_y"])


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_float_golden_bit_exact(kind, golden):
    if not Oracle.available(kind):
        pytest.skip(f"{kind} library not built")
    orc, g = Oracle(kind), golden["shift_golden"]
    names = g["names"].tolist()
    assert len(names) >= 240
    for name in names:
        dim, pad, active, borders = _case_args(name)
        x, w, grad = g[name + "/x"], g[name + "/w"], g[name + "/g"]
        y = orc.forward(x, w, pad, active, borders)
        assert np.array_equal(y, g[name + "/y"]), f"forward {name}"
        gi, gw = orc.backward(grad, x, w, pad, active, borders)
        assert np.array_equal(gi, g[name + "/gi"]), f"grad_input {name}"
        assert np.array_equal(gw, g[name + "/gw"]), f"grad_weight {name}"


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_channels_last_strides(kind, golden):
    """Strided (NHWC) input through the stride-aware fetch gives the NCHW answer."""
    if not Oracle.available(kind):
        pytest.skip(f"{kind} library not built")
    orc, g = Oracle(kind), golden["shift_golden"]
    for name in [n for n in g["names"].tolist() if n.startswith("cl_")]:
        dim, pad, active, _ = _case_args(name)
        x = g[name + "/x"]
        perm = (0, 2, 3, 1) if dim == 2 else (0, 2, 3, 4, 1)
        inv = (0, 3, 1, 2) if dim == 2 else (0, 4, 1, 2, 3)
        x_cl = np.ascontiguousarray(x.transpose(perm)).transpose(inv)     # NCHW view over NHWC memory
        assert not x_cl.flags["C_CONTIGUOUS"]
        assert np.array_equal(orc.forward(x_cl, g[name + "/w"], pad, active), g[name + "/y"]), name
        gi, gw = orc.backward(g[name + "/g"], x_cl, g[name + "/w"], pad, active)
        assert np.array_equal(gi, g[name + "/gi"]), name
        assert np.allclose(gw, g[name + "/gw"], rtol=1e-5, atol=1e-5), name   # NHWC loop order differs in the reference


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_quantized_golden_bit_exact(kind, golden):
    if not Oracle.available(kind):
        pytest.skip(f"{kind} library not built")
    orc, g = Oracle(kind), golden["quant_golden"]
    names = g["names"].tolist()
    assert len(names) >= 200
    for name in names:
        parts = name.split("_")
        dim, use_b, pad = int(parts[1][1]), int(parts[3][1]), int(parts[4][1])
        wzp, zp = g[name + "/meta"].tolist()
        y = orc.qforward(g[name + "/x"], g[name + "/wq"], wzp, zp, pad, BORDERS[dim] if use_b else None)
        assert y.dtype == g[name + "/y"].dtype and np.array_equal(y, g[name + "/y"]), name
        # the numpy restatement of quantize_shift_weights reproduces the reference's integer weights
        raw, z = quantize_shift_weights_np(g[name + "/wf"])
        assert z == wzp and np.array_equal(raw, g[name + "/wq"].astype(np.int64)), name


def test_port_equals_reference_headers_randomised(oracle_port, oracle_ref):
    """Wider randomised sweep: restatement == reference headers, all dims/pads/modes, odd sizes, borders."""
    rng = np.random.default_rng(7)
    shapes = {1: [(1, 2, 1), (2, 3, 2), (1, 4, 17)], 2: [(1, 2, 1, 7), (2, 2, 6, 1), (1, 3, 2, 2), (2, 2, 7, 9)],
              3: [(1, 2, 1, 1, 5), (1, 2, 3, 4, 5), (2, 1, 2, 2, 2)]}
    for dtype in (np.float32, np.float64):
        for dim, lst in shapes.items():
            for shape in lst:
                x = rng.standard_normal(shape).astype(dtype)
                w = ((rng.random((shape[1], dim)) * 2 - 1) * 6).astype(dtype)
                w.flat[0] = 0.5
                for borders in (None, [[1, 0]] * dim, [[0, 1]] * dim):
                    for pad in range(5):
                        for active in (False, True):
                            y1 = oracle_port.forward(x, w, pad, active, borders)
                            y2 = oracle_ref.forward(x, w, pad, active, borders)
                            assert np.array_equal(y1, y2), (shape, pad, active, borders)
                            grad = rng.standard_normal(y1.shape).astype(dtype)
                            a, b = oracle_port.backward(grad, x, w, pad, active, borders)
                            c, d = oracle_ref.backward(grad, x, w, pad, active, borders)
                            assert np.array_equal(a, c) and np.array_equal(b, d), (shape, pad, active, borders)


def test_reference_driver_threads_are_race_free(oracle_ref):
    from oracle.oracle import Oracle as O
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 6, 12, 10)).astype(np.float32)
    w = ((rng.random((6, 2)) * 2 - 1) * 2).astype(np.float32)
    g = rng.standard_normal(x.shape).astype(np.float32)
    gi1, gw1 = O("reference", threads=1).backward(g, x, w, 0, True)
    gi4, gw4 = O("reference", threads=4).backward(g, x, w, 0, True)
    gi4b, gw4b = O("reference", threads=4).backward(g, x, w, 0, True)
    assert np.array_equal(gi1, gi4) and np.array_equal(gw4, gw4b)
    assert np.allclose(gw1, gw4, rtol=1e-4, atol=1e-4)


def test_weight_split_matches_c_restatement(oracle_port):
    import ctypes as ct
    rng = np.random.default_rng(5)
    for dtype, sfx in ((np.float32, "f32"), (np.float64, "f64")):
        w = np.concatenate([(rng.random(200) * 2 - 1) * 9, [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.0, -0.0, 3.0, -3.0, -1e-10, 1e-10]]).astype(dtype)
        for active in (False, True):
            for which, fn_np in (("forward", split_forward), ("backward", split_backward)):
                iw = np.empty(w.shape, dtype=np.int64)
                dw = np.empty(w.shape, dtype=dtype)
                fn = getattr(oracle_port.lib, f"oracle_split_{which}_{sfx}")
                fn.restype = None
                fn(ct.c_int(int(active)), ct.c_int64(w.size), w.ctypes.data_as(ct.c_void_p),
                   iw.ctypes.data_as(ct.c_void_p), dw.ctypes.data_as(ct.c_void_p))
                iw2, dw2 = fn_np(w, active)
                assert np.array_equal(iw, iw2) and np.array_equal(dw, dw2), (sfx, active, which)


def test_check_borders_matches_reference_shapes(golden):
    for row in golden["kat"]["border_cases"].tolist():
        user = [[row[0], row[1]], [row[2], row[3]]]
        lb, rb = check_borders(2, (16, 16), user)
        assert [rb[0] - lb[0], rb[1] - lb[1]] == row[4:], row
    with pytest.raises(RuntimeError):
        check_borders(2, (16, 16), [[20, 0], [0, 0]])     # the reference dies with "negative dimension"


def test_remap_properties(oracle_port):
    idx = np.arange(-40, 40)
    for n in (2, 3, 5, 8):
        for pad in (1, 2, 3, 4):
            out = oracle_port.remap_axis(pad, n, idx)
            assert out.min() >= 0 and out.max() <= n - 1
            inside = (idx >= 0) & (idx < n)
            assert np.array_equal(out[inside], idx[inside])          # identity inside the axis
        z = oracle_port.remap_axis(0, n, idx)
        assert np.array_equal(z[idx < 0], idx[idx < 0]) and (z[idx >= n] == -1).all()
        per = oracle_port.remap_axis(2, n, idx)
        assert np.array_equal(per, np.mod(idx, n))
        sym = oracle_port.remap_axis(4, n, idx)
        assert np.array_equal(sym, np.where((np.floor_divide(idx, n) % 2) == 0, np.mod(idx, n), n - 1 - np.mod(idx, n)))
