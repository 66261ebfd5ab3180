#!/usr/bin/env python
"""Host-copy ceiling of the box for N ranks: every rank moves a cfg3-sized tensor pinned-host -> device and device ->
pinned-host concurrently (two streams), all ranks at once; prints per-rank and aggregate GB/s per direction and the
NUMA placement it ran with.  The bench's e2e pipeline moves 2 x 0.82 GB each way per GPU per step, so its ceiling is
this number.   torchrun --nproc-per-node N tools/e2e_ceiling_probe.py      (or plain python for one GPU)"""
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 256 * 256 * 56 * 56
a = torch.empty(n, pin_memory=True); b = torch.empty(n, pin_memory=True)
da = torch.empty(n, device=dev); db = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(both, reps=4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                b.copy_(db, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


run(True, 1)
for label, both in (("H2D only", False), ("H2D + D2H concurrently", True)):
    dt = run(both)
    if rank == 0:
        gb = n * 4 / dt / 1e9
        print(f"{world} rank(s), {label}: {gb:.1f} GB/s per direction per rank (slowest rank), {gb * world:.1f} GB/s aggregate per direction", flush=True)
if rank == 0:
    try:
        aff = sorted(os.sched_getaffinity(0))
        print(f"rank 0 CPU affinity: {aff[0]}-{aff[-1]} ({len(aff)} cpus); cpu_count {os.cpu_count()}")
    except Exception:
        pass
if world > 1:
    dist.destroy_process_group()
