#!/usr/bin/env python
"""Multi-rank check of the in-kernel grad_weight all-reduce (ts_shift_backward_allreduce) against
torch.distributed's NCCL all-reduce, and its latency.  Run under torchrun on 2..8 GPUs:
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fused_allreduce_probe.py
Checks (every one against NCCL's sum of the plain backward, rtol 1e-5, and bit-identical across ranks):
  1. eager steps over several layer shapes / paddings / dims (the per-CTA device call counters advance
     differently for layers with different C x dim);
  2. a RAGGED batch whose last ranks get an EMPTY shard (they must still take part and contribute zeros);
  3. the step captured in CUDA graphs (GraphedShiftStep) and replayed on new data;
  4. latency of a 32-image-per-GPU step: eager NCCL, eager in-kernel, graphed in-kernel.
`tests/test_gpu_parity.py::test_fused_allreduce_two_ranks` runs this file with --quick on 2 GPUs."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")]
import torchshifts  # noqa: E402,F401
from torchshifts.functional import shift1d_func, shift2d_func, shift3d_func  # noqa: E402
from torchshifts.host import GraphedShiftStep  # noqa: E402
from torchshifts.sharded import FusedGradWeightAllReduce, shard_range  # noqa: E402

quick = "--quick" in sys.argv
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
fused = FusedGradWeightAllReduce(capacity=4096, device=dev)
FN = {1: shift1d_func, 2: shift2d_func, 3: shift3d_func}
ok = True


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def compare(tag, got, ref):
    global ok
    err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    same = [torch.empty_like(got) for _ in range(world)]
    dist.all_gather(same, got.contiguous())
    identical = all(torch.equal(same[0], t) for t in same)
    say(f"{tag}: max rel err vs NCCL {err:.2e}; identical on all ranks: {identical}")
    ok = ok and err < 1e-5 and identical


# ---- 1. eager, mixed layers --------------------------------------------------------------------
cases = [((8, 64, 32, 32), 0, False), ((8, 64, 32, 32), 3, True), ((16, 256, 56, 56), 0, False), ((4, 16, 8, 12, 16), 2, True),
         ((6, 1000, 64), 1, True), ((3, 2048, 4, 8), 4, False)]
for it, (shape, pad, active) in enumerate(cases * (1 if quick else 3)):
    dim = len(shape) - 2
    torch.manual_seed(100 * it + rank)
    x = torch.randn(shape, device=dev, requires_grad=True)
    g = torch.randn(shape, device=dev)
    torch.manual_seed(it)
    w = ((torch.rand(shape[1], dim, device=dev) * 2 - 1) * 2).requires_grad_(True)
    FN[dim](x, w, pad, active).backward(g)
    ref = w.grad.clone()
    dist.all_reduce(ref)
    x.grad = None; w.grad = None
    with fused:
        FN[dim](x, w, pad, active).backward(g)
    compare(f"eager case {it} {shape} pad {pad} active {active}", w.grad, ref)

# ---- 2. ragged batch: ranks beyond the batch own an EMPTY shard ---------------------------------
total = max(1, world - 1)            # one rank (at least) gets nothing
torch.manual_seed(7)
xa = torch.randn(total, 32, 16, 16, device=dev)
ga = torch.randn(total, 32, 16, 16, device=dev)
wa = (torch.rand(32, 2, device=dev) * 2 - 1)
lo, hi = shard_range(total, rank, world)
xs, gs = xa[lo:hi].clone().requires_grad_(True), ga[lo:hi].clone()
ws = wa.clone().requires_grad_(True)
xf, wf = xa.clone().requires_grad_(True), wa.clone().requires_grad_(True)
shift2d_func(xf, wf, 3, True).backward(ga)                     # full batch on every rank = the expected sum
with fused:
    shift2d_func(xs, ws, 3, True).backward(gs)
compare(f"ragged batch of {total} over {world} ranks (rank {world - 1} empty)", ws.grad, wf.grad)

# ---- 3. graph capture + replay -------------------------------------------------------------------
shape = (8, 64, 32, 32) if quick else (32, 256, 56, 56)
torch.manual_seed(rank)
x = torch.randn(shape, device=dev); g = torch.randn(shape, device=dev)
torch.manual_seed(0)
w = torch.rand(shape[1], 2, device=dev) * 2 - 1
with fused:
    graphed = GraphedShiftStep(x, w, g, 0, False, split=True)
for rep in range(3):
    torch.manual_seed(50 + 10 * rep + rank)
    graphed.x.detach().copy_(torch.randn(shape, device=dev)); graphed.grad_out.copy_(torch.randn(shape, device=dev))
    xr, wr = graphed.x.detach().clone().requires_grad_(True), w.clone().requires_grad_(True)
    shift2d_func(xr, wr, 0, False).backward(graphed.grad_out)
    ref = wr.grad.clone()
    dist.all_reduce(ref)
    y, gi, gw = graphed.replay()
    torch.cuda.synchronize()
    compare(f"graph replay {rep}", gw, ref)
    ok = ok and torch.equal(gi, xr.grad)


# ---- 4. latency -----------------------------------------------------------------------------------
def timed(body, reps=30):
    for _ in range(5):
        body()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        body()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1000], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if not quick:
    xx = graphed.x.detach().clone().requires_grad_(True)
    ww = w.clone().requires_grad_(True)

    def step_nccl():
        xx.grad = None; ww.grad = None
        shift2d_func(xx, ww, 0, False).backward(g)
        dist.all_reduce(ww.grad)

    def step_fused():
        xx.grad = None; ww.grad = None
        with fused:
            shift2d_func(xx, ww, 0, False).backward(g)

    plain = GraphedShiftStep(x, w, g, 0, False, split=True)           # no exchange at all: the device-work floor
    t_n, t_f, t_g, t_p = timed(step_nccl), timed(step_fused), timed(graphed.replay), timed(plain.replay)
    say(f"fwd+bwd on N={shape[0]} per GPU, {world} GPUs (max over ranks): eager + NCCL all-reduce {t_n:.1f} us/step, eager + in-kernel "
        f"exchange {t_f:.1f} us/step, graphed + in-kernel exchange {t_g:.1f} us/step, graphed without any exchange {t_p:.1f} us/step")
say("PROBE", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
