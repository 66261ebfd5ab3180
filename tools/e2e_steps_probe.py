#!/usr/bin/env python
"""Per-step wall time of the host-buffer pipeline (bench.py's e2e leg) from a cold start: shows how many steps a fresh box
needs before the pinned-copy pipeline reaches its steady rate.   python tools/e2e_steps_probe.py [steps]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.host import HostShift2dPipeline  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
pipe = HostShift2dPipeline(256, 256, 56, 56, device=dev)
pipe.x_host.normal_(); pipe.g_host.normal_()
w = torch.rand(256, 2, device=dev) * 2 - 1
ts = []
for i in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gw = pipe.forward_backward(w, 0, False)
    pipe.read_back_grad_weight(gw)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("ms per step:", " ".join(f"{t:.1f}" for t in ts))
