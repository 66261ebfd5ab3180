"""Quantized functional interface (reference: ``torchshifts/quantized/functional.py``): the input
must be a quantized tensor, the shift is always the sparse (integer) one."""
from torchshifts.functional import shift1d_func, shift2d_func, shift3d_func

_FUNCS = {1: shift1d_func, 2: shift2d_func, 3: shift3d_func}


def _quantized(dim, input, weight, padding_mode, cut_borders):
    if not input.is_quantized:
        raise ValueError(f"Input to 'shift{dim}d_quantized' must be quantized!")
    return _FUNCS[dim](input, weight, padding_mode, False, cut_borders)


def shift1d_quantized(input, weight, padding_mode, cut_borders=None):
    return _quantized(1, input, weight, padding_mode, cut_borders)


def shift2d_quantized(input, weight, padding_mode, cut_borders=None):
    return _quantized(2, input, weight, padding_mode, cut_borders)


def shift3d_quantized(input, weight, padding_mode, cut_borders=None):
    return _quantized(3, input, weight, padding_mode, cut_borders)
