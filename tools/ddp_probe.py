#!/usr/bin/env python
"""Which part of a sharded fwd+bwd+all-reduce step costs what (diagnostic for bench.py --gpus N)."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.functional import shift2d_func  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(N, 256, 56, 56, device=dev).requires_grad_(True)
g = torch.randn(N, 256, 56, 56, device=dev)
w = (torch.rand(256, 2, device=dev) * 2 - 1).requires_grad_(True)
buf = torch.zeros(512, device=dev)
fwd = torch.ops.torchshifts._shift2d_forward
bwd = torch.ops.torchshifts._shift2d_backward
borders = torch.tensor([0, 56, 0, 56, 0, 1], dtype=torch.int32)


def timed(label, body, reps=20):
    for _ in range(3):
        body()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        body()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{label:50s} device {a.elapsed_time(b) / reps * 1000:9.1f} us/step | host enqueue {(t1 - t0) / reps * 1e6:9.1f} us/step", flush=True)


def autograd_step(reduce):
    def f():
        x.grad = None; w.grad = None
        y = shift2d_func(x, w, 0, False)
        y.backward(g)
        if reduce == "grad" and world > 1:
            dist.all_reduce(w.grad)
        elif reduce == "buf" and world > 1:
            dist.all_reduce(buf)
    return f


def raw_step(reduce):
    def f():
        with torch.no_grad():
            y = fwd(x, w, borders, list(x.shape), 0, False)
            gi, gw = bwd(g, w, x, borders, 0, False)
            if reduce and world > 1:
                dist.all_reduce(gw)
    return f


timed("autograd fwd+bwd, no all_reduce", autograd_step(None))
timed("autograd fwd+bwd + all_reduce(separate buffer)", autograd_step("buf"))
timed("autograd fwd+bwd + all_reduce(w.grad)", autograd_step("grad"))
timed("raw ops fwd+bwd, no all_reduce", raw_step(False))
timed("raw ops fwd+bwd + all_reduce(gw)", raw_step(True))
if rank == 0:
    print(torch.cuda.memory_summary(abbreviated=True)[:1800])
if world > 1:
    dist.destroy_process_group()
