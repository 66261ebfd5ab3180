#!/usr/bin/env python
"""Does the pinned-copy rate depend on the allocation?  Repeatedly allocates two pinned 0.8 GB buffers, measures H2D + D2H
running concurrently, frees them; prints the rate of every round plus the NUMA facts of the box."""
import glob
import os
import time

import torch

dev = torch.device("cuda:0")
n = 256 * 256 * 56 * 56
da = torch.empty(n, device=dev); db = torch.empty(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
print("nodes:", [os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node*")], "affinity:", sorted(os.sched_getaffinity(0)))
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.dirname(p) + "/class").read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print("gpu", os.path.dirname(p).split("/")[-1], "numa_node", open(p).read().strip())
    except OSError:
        pass
try:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
except OSError:
    pass
for r in range(8):
    a = torch.empty(n, pin_memory=True); b = torch.empty(n, pin_memory=True)
    a.fill_(1.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            da.copy_(a, non_blocking=True)
        with torch.cuda.stream(s2):
            b.copy_(db, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"round {r}: {n * 4 / dt / 1e9:.1f} GB/s per direction (both directions at once)", flush=True)
    del a, b
    torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None
