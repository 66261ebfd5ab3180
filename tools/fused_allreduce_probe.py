#!/usr/bin/env python
"""Multi-rank check of the in-kernel grad_weight all-reduce (ts_shift_backward_allreduce) against
torch.distributed's NCCL all-reduce, and its latency.  Run under torchrun on 2..8 GPUs:
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_allreduce_probe.py"""
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
from torchshifts.sharded import FusedGradWeightAllReduce  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
fused = FusedGradWeightAllReduce(capacity=4096, device=dev)
ok = True
for it, (shape, pad, active) in enumerate([((8, 64, 32, 32), 0, False), ((8, 64, 32, 32), 3, True), ((16, 256, 56, 56), 0, False)] * 3):
    torch.manual_seed(100 * it + rank)
    x = torch.randn(shape, device=dev, requires_grad=True)
    g = torch.randn(shape, device=dev)
    torch.manual_seed(it)
    w = (torch.rand(shape[1], 2, device=dev) * 2 - 1).requires_grad_(True)
    shift2d_func(x, w, pad, active).backward(g)
    ref = w.grad.clone()
    dist.all_reduce(ref)
    x.grad = None; w.grad = None
    with fused:
        shift2d_func(x, w, pad, active).backward(g)
    got = w.grad
    err = float((got - ref).abs().max() / ref.abs().max())
    same = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(same, got)
    identical = all(torch.equal(same[0], t) for t in same)
    if rank == 0:
        print(f"case {it}: max rel err vs NCCL {err:.2e}; identical on all ranks: {identical}", flush=True)
    ok = ok and err < 1e-5 and identical
# latency: backward + reduction, fused vs NCCL
shape = (32, 256, 56, 56)
x = torch.randn(shape, device=dev, requires_grad=True); g = torch.randn(shape, device=dev)
w = (torch.rand(256, 2, device=dev) * 2 - 1).requires_grad_(True)


def timed(body, reps=30):
    for _ in range(5):
        body()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        body()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1000


def step_nccl():
    x.grad = None; w.grad = None
    shift2d_func(x, w, 0, False).backward(g)
    dist.all_reduce(w.grad)


def step_fused():
    x.grad = None; w.grad = None
    with fused:
        shift2d_func(x, w, 0, False).backward(g)


t_n = timed(step_nccl); t_f = timed(step_fused)
if rank == 0:
    print(f"fwd+bwd on N=32 per GPU, {world} GPUs: NCCL all-reduce {t_n:.1f} us/step, in-kernel exchange {t_f:.1f} us/step", flush=True)
    print("PROBE", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
