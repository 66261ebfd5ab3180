"""torchshifts (B200 build): drop-in for the Sparse/Active Shift operator of
DeadAt0m/ActiveSparseShifts-PyTorch with the compute in hand-written sm_100a CUDA kernels behind a
C ABI (``libtorchshifts_b200.so``).  Public surface as in the reference's ``torchshifts/__init__.py``."""
from .extension import _HAS_OPS
from .version import __version__

from torchshifts.modules.shifts import Shift1d, Shift2d, Shift3d
from torchshifts.quantized import quant_mapping

__all__ = ['Shift1d', 'Shift2d', 'Shift3d', 'quant_mapping', '__version__']
