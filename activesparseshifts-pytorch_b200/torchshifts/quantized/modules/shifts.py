"""Quantized shift layers (reference: ``torchshifts/quantized/modules/shifts.py``).

``quantize_shift_weights`` keeps the reference's convention exactly: ``scale = ceil((max-min)/255)``,
zero point 128, quint8 -- so the effective integer shift is ``round_half_even(w / scale)`` and the
scale is never multiplied back.  Modules return a bare tensor (no loss).
"""
import math

import torch

import torchshifts.modules.shifts as shifts
from torchshifts.quantized.functional import shift1d_quantized, shift2d_quantized, shift3d_quantized

rp_dict = {v: k for k, v in shifts.paddings_dict.items()}
_QFUNCS = {1: shift1d_quantized, 2: shift2d_quantized, 3: shift3d_quantized}


def quantize_shift_weights(weight):
    scale = math.ceil((weight.max().item() - weight.min().item()) / 255.)
    return torch.quantize_per_tensor(weight, scale, 128, torch.quint8)


def _make_quantized(dim, base):
    class _QShift(base):
        def __init__(self, in_channels, padding='zeros'):
            super().__init__(in_channels, padding, 1, 0, False)
            self.qweight = quantize_shift_weights(self.weight.float())

        def forward(self, input):
            qweight = self.qweight
            if qweight.device != input.device:      # qweight is a plain attribute: .to(device) does not move it
                qweight = self.qweight = qweight.to(input.device)
            return self._reduction_fn(_QFUNCS[dim](input, qweight, self.padding, self.cut_borders))

        def _get_name(self):
            return f'QuantizedShift{dim}D'

        @staticmethod
        def from_float(mod):
            qshift = _QShift(mod.in_channels, rp_dict[mod.padding])
            qshift.cut_borders = mod.cut_borders
            qshift._reduction_fn = mod._reduction_fn
            qshift.weight = mod.weight
            qshift.qweight = quantize_shift_weights(mod.weight.float())
            return qshift

    _QShift.__name__ = _QShift.__qualname__ = f'Shift{dim}d'
    _QShift.__doc__ = f'Quantized counterpart of :class:`torchshifts.modules.shifts.Shift{dim}d`; built by ``from_float``.'
    return _QShift


Shift1d = _make_quantized(1, shifts.Shift1d)
Shift2d = _make_quantized(2, shifts.Shift2d)
Shift3d = _make_quantized(3, shifts.Shift3d)
