#!/usr/bin/env python
"""Latency of the path's one collective (all-reduce of a [C,2] fp32 tensor) under torchrun, alone and
interleaved with a bandwidth-bound kernel, in several variants (diagnostic for bench.py --gpus N)."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
t = torch.ones(512, device=dev)
for _ in range(5):
    dist.all_reduce(t)
torch.cuda.synchronize()
x = torch.randn(64, 256, 56, 56, device=dev)
y = torch.empty_like(x)


def timed(label, body, reps=20):
    for _ in range(3):
        body()
    torch.cuda.synchronize()
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
    t2 = time.perf_counter()
    if rank == 0:
        print(f"{label:44s} device {a.elapsed_time(b) / reps * 1000:8.1f} us/iter | host enqueue {(t1 - t0) / reps * 1e6:8.1f} us/iter | "
              f"wall {(t2 - t0) / reps * 1e6:8.1f} us/iter", flush=True)


def ar():
    dist.all_reduce(t)


def mul():
    torch.mul(x, 2.0, out=y)


def mul_ar():
    torch.mul(x, 2.0, out=y)
    dist.all_reduce(t)


works = []


def mul_ar_async():
    torch.mul(x, 2.0, out=y)
    works.append(dist.all_reduce(t, async_op=True))
    if len(works) > 4:
        works.pop(0).wait()


side = torch.cuda.Stream(dev)


def mul_side():
    torch.mul(x, 2.0, out=y)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        t.add_(1.0)
    torch.cuda.current_stream().wait_stream(side)


timed("all_reduce(512 floats)", ar, 50)
timed("mul (411 MB of traffic)", mul)
timed("mul + all_reduce", mul_ar)
timed("mul + all_reduce(async_op, wait 4 later)", mul_ar_async)
timed("mul + tiny kernel on a side stream", mul_side)
os.environ["X"] = "1"
timed("mul + all_reduce (again)", mul_ar)
dist.destroy_process_group()
