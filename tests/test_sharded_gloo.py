"""world_size-2 gloo test of the N>1 path (host-side logic of torchshifts.sharded).

There is no GPU here and the product has no CPU compute path, so each rank computes its shard's
forward / backward with the ORACLE (test infrastructure) and the test checks what the sharding
layer is responsible for: shards tile the batch, forward and grad_input need no exchange, and one
all-reduce of the coalesced grad_weight buffer reproduces the full-batch grad_weight."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        from torchshifts import Shift1d, Shift2d
        from torchshifts.sharded import allreduce_grad_weights, broadcast_weights, shard_batch, shard_range
        orc = Oracle("port")
        torch.manual_seed(100 + rank)                       # deliberately different init per rank
        layers = torch.nn.ModuleList([Shift2d(5, init_shift=2), Shift1d(3, init_shift=3), Shift2d(2)])
        broadcast_weights(layers, src=0)
        w_all = [m.weight.detach().clone() for m in layers]
        gathered = [torch.zeros_like(w_all[0]) for _ in range(world)]
        dist.all_gather(gathered, w_all[0])
        assert all(torch.equal(gathered[0], t) for t in gathered), "weights not replicated"
        rng = np.random.default_rng(42)                      # same data on every rank, then sharded
        N = 7                                                # ragged: 4 + 3
        data = [(rng.standard_normal((N, 5, 6, 8)).astype(np.float32), 0, False),
                (rng.standard_normal((N, 3, 16)).astype(np.float32), 2, True),
                (rng.standard_normal((N, 2, 4, 4)).astype(np.float32), 3, True)]
        grads = [rng.standard_normal(d[0].shape).astype(np.float32) for d in data]
        lo, hi = shard_range(N, rank, world)
        assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))
        local_out = []
        for m, (x, pad, active), g in zip(layers, data, grads):
            xs = shard_batch(torch.from_numpy(x), rank, world).numpy()
            gs = g[lo:hi]
            w = m.weight.detach().numpy()
            y = orc.forward(xs, w, pad, active)
            gi, gw = orc.backward(gs, xs, w, pad, active)
            m.weight.grad = torch.from_numpy(gw)
            local_out.append((y, gi))
        n = allreduce_grad_weights(layers)
        assert n == 5 * 2 + 3 * 1 + 2 * 2
        # full-batch truth, computed redundantly on every rank
        for m, (x, pad, active), g, (y, gi) in zip(layers, data, grads, local_out):
            w = m.weight.detach().numpy()
            y_full = orc.forward(x, w, pad, active)
            gi_full, _ = orc.backward(g, x, w, pad, active)
            _, gw64 = orc.backward(g.astype(np.float64), x.astype(np.float64), w.astype(np.float64), pad, active)
            assert np.array_equal(y, y_full[lo:hi]) and np.array_equal(gi, gi_full[lo:hi])      # no exchange needed
            assert np.allclose(m.weight.grad.numpy(), gw64, rtol=1e-5, atol=1e-5 * np.abs(gw64).max())
        # a rank with an empty shard still takes part in the collective
        assert shard_range(1, 1, 2) == (1, 1) and shard_range(10, 3, 4) == (8, 10)
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_batch_sharding_and_single_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_range_tiles_the_batch():
    sys.path.insert(0, str(ROOT / "activesparseshifts-pytorch_b200"))
    from torchshifts.sharded import shard_range
    for n in (0, 1, 7, 8, 255, 256):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def _ddp_worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(ROOT / "activesparseshifts-pytorch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torchshifts import Shift2d
        from torchshifts.sharded import FusedGradWeightAllReduce, ddp_sum_to_mean_hook

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.lin = torch.nn.Linear(4, 4)
                self.shift = Shift2d(3, init_shift=2)          # never run here: there is no CPU compute path
                self.block = torch.nn.Sequential(Shift2d(2))

            def forward(self, x):
                return self.lin(x)

        torch.manual_seed(0)
        net = Net()
        names = FusedGradWeightAllReduce.exclude_from_ddp(net)
        assert sorted(names) == ["block.0.weight", "shift.weight"]
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        # DDP keeps only the Linear's parameters in its buckets: the shift weights are reduced by the in-kernel exchange
        managed = {n for n, p in ddp.module.named_parameters() if n not in ddp.parameters_to_ignore}
        assert managed == {"lin.weight", "lin.bias"}, managed
        x = torch.full((2, 4), float(rank + 1))
        ddp(x).sum().backward()
        mean_x = sum(range(1, world + 1)) / world                 # DDP averaged the Linear's gradient over the ranks
        assert torch.allclose(net.lin.weight.grad, torch.full((4, 4), 2 * mean_x))
        # the fused exchange SUMS; the hook turns that into DDP's mean for the excluded weights
        handles = ddp_sum_to_mean_hook(net)
        assert len(handles) == 2
        (net.shift.weight * 3.0).sum().backward()                 # stands for "grad already summed over the ranks"
        assert torch.allclose(net.shift.weight.grad, torch.full_like(net.shift.weight, 3.0 / world))
        Path(out_dir, f"ddp{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_ddp_leaves_the_shift_weights_to_the_fused_exchange(tmp_path):
    """with the in-kernel exchange the shift weights must NOT sit in DDP's gradient buckets as well
    (they would be reduced twice): exclude_from_ddp + the sum->mean hook, world size 2 on gloo."""
    world, port = 2, _free_port()
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ddp{r}").exists() for r in range(world))
