import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts
from torchshifts.extension import native
from torchshifts.functional import shift2d_func
lib = native().lib
dev = torch.device("cuda:0")
rng = np.random.default_rng(61)
for shape, pad_name, active, conv_pad in [((2, 8, 16, 16), 'zeros', False, 1), ((2, 8, 15, 20), 'reflect', True, 1), ((3, 4, 18, 26), 'periodic', False, 0),
                                          ((2, 4, 17, 14), 'symmetric', True, 0), ((2, 6, 16, 24), 'border', True, 1), ((2, 4, 9, 22), 'zeros', True, 1)]:
    torch.manual_seed(3)
    m = torchshifts.Shift2d(shape[1], padding=pad_name, active_flag=active, sparsity_term=0, emulate_dw={'kernel_size': 3, 'stride': 2, 'padding': conv_pad}).to(dev)
    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).to(dev)
    out, loss = m(x)
    path = lib.ts_last_kernel_path()
    pad = torchshifts.modules.shifts.paddings_dict[pad_name]
    two = torch.nn.functional.avg_pool2d(shift2d_func(x, m.weight, pad, active, m.cut_borders), 2, 2, ceil_mode=True)
    print(shape, pad_name, active, conv_pad, "path", path, "fused?", m._pool_is_fused(x), "stride", m._stride_ints, "out", tuple(out.shape), "two", tuple(two.shape),
          "equal", torch.equal(out, two), "maxdiff", float((out - two).abs().max()), "w range", float(m.weight.min()), float(m.weight.max()))
