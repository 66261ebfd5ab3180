from .shifts import Shift1d, Shift2d, Shift3d
import torchshifts.modules.shifts as _float_shifts

# float module -> quantized module, for torch.quantization.convert(model, mapping=quant_mapping)
new_quant_mapping = {_float_shifts.Shift1d: Shift1d, _float_shifts.Shift2d: Shift2d, _float_shifts.Shift3d: Shift3d}
