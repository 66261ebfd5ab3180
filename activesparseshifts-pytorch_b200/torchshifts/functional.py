"""Functional interface: ``shift{1,2,3}d_func`` -- same names, argument order, argument checks and
error behaviour as the reference's ``torchshifts/functional.py:7-99`` (bare ``assert`` with the
reference's messages, ``borders=None`` becomes an empty tensor, then the registered operator
``torch.ops.torchshifts.shift{1,2,3}d`` is called).  The compute behind the operator is the
sm_100a library; tensors must live on a CUDA device.
"""
from typing import Optional

import torch

from .extension import _assert_has_ops

Tensor = torch.Tensor
_AXES = {1: 'H', 2: 'H and W', 3: 'H,W and D'}
_MODES = '0 - zeros, 1 - border, 2 - periodic, 3 - reflect, 4 - symmetric'


def _shift_func(dim: int, input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                borders: Optional[Tensor]) -> Tensor:
    name = f'shift{dim}d_func()'
    _assert_has_ops()
    assert padding_mode in [0, 1, 2, 3, 4], f'{name} expected padding_mode can be {_MODES}'
    assert len(input.shape) == dim + 2, f'{name[:-2]}(): expected {dim + 2}D tensor as input, but it is shape is {input.shape}'
    assert weights.shape[-1] == dim, f'{name[:-2]}(): expected [n_channels,{dim}] tensor as weight, but it is shape is {weights.shape}'
    assert input.shape[1] == weights.shape[0], (f'{name[:-2]}(): expected that input and weight have equal number of channels, '
                                                f'but input have {input.shape[1]} and weight have {weights.shape[0]} channels.')
    assert input.device == weights.device, (f'{name[:-2]}(): expected input and weights to be on same device, '
                                            f'but input is  on {input.device} and weights is on {weights.device}')
    if borders is not None:
        assert (len(borders.shape) == 2) and (borders.shape[1] == 2) and (borders.shape[0] == dim), f'borders must have shape [{dim}, 2]'
    else:
        borders = torch.Tensor()
    op = getattr(torch.ops.torchshifts, f'shift{dim}d')
    return op(input, weights, borders, padding_mode, active_flag)


def shift1d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None) -> Tensor:
    """Shift every channel of ``input [N, C, H]`` along H by its own learnable amount.

    ``weights [C, 1]`` holds sign and magnitude of the per-channel shift; ``padding_mode`` selects
    what is read outside the tensor (0 zeros, 1 border, 2 periodic, 3 reflect, 4 symmetric);
    ``active_flag`` switches from the rounded (sparse) shift to linear interpolation (ignored for
    quantized inputs); ``borders [1, 2]`` = (left_cut, right_cut) crops the output.
    """
    return _shift_func(1, input, weights, padding_mode, active_flag, borders)


def shift2d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None) -> Tensor:
    """2-D version: ``input [N, C, H, W]``, ``weights [C, 2]`` (H and W shifts), ``borders [2, 2]``;
    the active shift is bilinear."""
    return _shift_func(2, input, weights, padding_mode, active_flag, borders)


def shift3d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None) -> Tensor:
    """3-D version: ``input [N, C, H, W, D]``, ``weights [C, 3]``, ``borders [3, 2]``; the active
    shift is trilinear."""
    return _shift_func(3, input, weights, padding_mode, active_flag, borders)


# BASELINE.json names these "functional.shift1d/2d/3d"; keep the real names and add the aliases.
shift1d, shift2d, shift3d = shift1d_func, shift2d_func, shift3d_func
