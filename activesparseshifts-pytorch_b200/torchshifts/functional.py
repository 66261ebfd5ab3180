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


def _resolve_borders(dim, sizes, cuts):
    """Pure-Python twin of ``ts_check_borders`` (csrc/ops/shifts.cpp:111-127): ``cuts`` = ``dim`` pairs
    (left_cut, right_cut) -> (lb, rb) lists of 3.  Used only while ``torch.compile`` traces a layer that crops
    its output: there the border tensor is a traced (fake) tensor whose values cannot be read, while the
    module's cached Python integers can (tests pin this function against the C one)."""
    lb, rb = [0, 0, 0], [int(sizes[a]) if a < dim else 1 for a in range(3)]
    for a in range(dim):
        size = int(sizes[a])
        r, l = size - int(cuts[a][1]), int(cuts[a][0])
        if r - l < 1:
            r = l + 1
        if l == size:
            l, r = size - 1, size
        if r == 0:
            l, r = 0, 1
        l = max(l, 0)
        r = min(r, size)
        if r - l < 0:
            raise RuntimeError('torchshifts-b200: borders give a negative output dimension '
                               '(the reference fails here with "Trying to create tensor with negative dimension")')
        lb[a], rb[a] = l, r
    return lb, rb


@torch.compiler.assume_constant_result
def _std_borders(vals):
    """int32[6] host tensor {l0,r0,l1,r1,l2,r2}; a constant of the compiled graph (no CPU code is generated for it)"""
    return torch.tensor(list(vals), dtype=torch.int32)


def _shift_func(dim: int, input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                borders: Optional[Tensor], _border_ints=None) -> Tensor:
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
        if _border_ints is not None and torch.compiler.is_compiling():
            # torch.compile: resolve the crop from Python integers (constants of the trace) and call the inner
            # operator, which has a fake kernel and a registered autograd formula
            lb, rb = _resolve_borders(dim, input.shape[2:], _border_ints)
            std = _std_borders((lb[0], rb[0], lb[1], rb[1], lb[2], rb[2]))
            new_size = [input.shape[0], input.shape[1]] + [rb[a] - lb[a] for a in range(dim)]
            return getattr(torch.ops.torchshifts, f'_shift{dim}d_forward')(input, weights, std, new_size, padding_mode, active_flag)
    else:
        borders = torch.Tensor()
    op = getattr(torch.ops.torchshifts, f'shift{dim}d')
    return op(input, weights, borders, padding_mode, active_flag)


def shift1d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None, _border_ints=None) -> Tensor:
    """Shift every channel of ``input [N, C, H]`` along H by its own learnable amount.

    ``weights [C, 1]`` holds sign and magnitude of the per-channel shift; ``padding_mode`` selects
    what is read outside the tensor (0 zeros, 1 border, 2 periodic, 3 reflect, 4 symmetric);
    ``active_flag`` switches from the rounded (sparse) shift to linear interpolation (ignored for
    quantized inputs); ``borders [1, 2]`` = (left_cut, right_cut) crops the output.
    """
    return _shift_func(1, input, weights, padding_mode, active_flag, borders, _border_ints)


def shift2d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None, _border_ints=None) -> Tensor:
    """2-D version: ``input [N, C, H, W]``, ``weights [C, 2]`` (H and W shifts), ``borders [2, 2]``;
    the active shift is bilinear."""
    return _shift_func(2, input, weights, padding_mode, active_flag, borders, _border_ints)


def shift3d_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                 borders: Optional[Tensor] = None, _border_ints=None) -> Tensor:
    """3-D version: ``input [N, C, H, W, D]``, ``weights [C, 3]``, ``borders [3, 2]``; the active
    shift is trilinear."""
    return _shift_func(3, input, weights, padding_mode, active_flag, borders, _border_ints)


def shift2d_avgpool2_func(input: Tensor, weights: Tensor, padding_mode: int, active_flag: bool,
                          borders: Optional[Tensor] = None, _border_ints=None) -> Tensor:
    """``avg_pool2d(shift2d_func(...), kernel_size=2, stride=2, ceil_mode=True)`` in one kernel: what a ``Shift2d`` built
    with ``emulate_dw={..., 'stride': 2}`` computes (reference ``modules/shifts.py:85-89, 153``), with one read of the
    input and one quarter-size write.  Same argument checks as :func:`shift2d_func`."""
    name = 'shift2d_avgpool2_func'
    _assert_has_ops()
    assert padding_mode in [0, 1, 2, 3, 4], f'{name}() expected padding_mode can be {_MODES}'
    assert len(input.shape) == 4, f'{name}(): expected 4D tensor as input, but it is shape is {input.shape}'
    assert weights.shape[-1] == 2, f'{name}(): expected [n_channels,2] tensor as weight, but it is shape is {weights.shape}'
    assert input.shape[1] == weights.shape[0], f'{name}(): expected that input and weight have equal number of channels'
    assert input.device == weights.device, f'{name}(): expected input and weights to be on same device'
    if borders is None:
        cuts = [[0, 0], [0, 0]]
    else:
        assert tuple(borders.shape) == (2, 2), 'borders must have shape [2, 2]'
        cuts = _border_ints if _border_ints is not None else [[int(v) for v in row] for row in borders.tolist()]
    lb, rb = _resolve_borders(2, input.shape[2:], cuts)
    std = _std_borders((lb[0], rb[0], lb[1], rb[1], lb[2], rb[2]))
    new_size = [input.shape[0], input.shape[1], rb[0] - lb[0], rb[1] - lb[1]]
    return torch.ops.torchshifts._shift2d_avgpool2_forward(input, weights, std, new_size, padding_mode, active_flag)


# BASELINE.json names these "functional.shift1d/2d/3d"; keep the real names and add the aliases.
shift1d, shift2d, shift3d = shift1d_func, shift2d_func, shift3d_func
