"""``torchshifts.quantized``: quantized modules and the ``quant_mapping`` to hand to
``torch.quantization.convert`` (reference: ``torchshifts/quantized/__init__.py``).  Only the public
mapping getter of torch is used, so this works on torch 2.x where the reference's
``DEFAULT_OP_LIST_TO_FUSER_METHOD`` lookup raises AttributeError."""
try:
    from torch.ao.quantization.quantization_mappings import get_default_static_quant_module_mappings
except ImportError:  # pragma: no cover - very old torch
    from torch.quantization.quantization_mappings import get_default_static_quant_module_mappings

from .modules import Shift1d, Shift2d, Shift3d, new_quant_mapping

quant_mapping = {**get_default_static_quant_module_mappings(), **new_quant_mapping}
