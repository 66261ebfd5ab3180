"""Batch-sharded execution across the GPUs of one box (net new, the reference has no
distributed code).

Every (n, c) plane is independent in the forward pass and in grad_input; only
``grad_weight[c, :] = sum over n and space`` couples the batch.  So: one process per GPU
(torchrun), rank r owns ``x[shard_range(N, r, G)]``, weights are replicated, there is NO data-path
collective in forward or grad_input, and exactly ONE all-reduce (sum) of the tiny ``C x dim``
grad_weight per backward -- for all shift layers of a model coalesced into a single flat buffer.
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import ctypes as ct
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of a batch of ``n`` for ``rank`` of ``world`` (first
    ``n % world`` ranks get one extra item; ranks beyond ``n`` get an empty slice)."""
    assert 0 <= rank < world
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def _shift_weights(module_or_params) -> Iterable[torch.nn.Parameter]:
    if isinstance(module_or_params, torch.nn.Module):
        from torchshifts.modules.shifts import _Shiftnd
        return [m.weight for m in module_or_params.modules() if isinstance(m, _Shiftnd)]
    return list(module_or_params)


def allreduce_grad_weights(module_or_params, group=None, average: bool = False) -> int:
    """Sum (or average) ``weight.grad`` of every shift layer over the ranks with ONE collective.

    Returns the number of elements reduced.  Parameters without a gradient contribute zeros so
    that all ranks issue the same collective."""
    params = [p for p in _shift_weights(module_or_params)]
    if not params:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        piece = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = piece.clone()
        else:
            p.grad.copy_(piece)
        off += n
    return int(flat.numel())


def broadcast_weights(module_or_params, src: int = 0, group=None) -> None:
    """Make the replicated shift weights identical on every rank (one broadcast)."""
    params = [p for p in _shift_weights(module_or_params)]
    if not params:
        return
    flat = torch.cat([p.data.reshape(-1).to(torch.float32) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for p in params:
        n = p.numel()
        p.data.copy_(flat[off:off + n].reshape(p.shape).to(p.dtype))
        off += n


class FusedGradWeightAllReduce:
    """grad_weight leaves the backward already summed over the ranks: the deterministic pass-2
    reduction kernel exchanges the ``C x dim`` values over NVLink peer memory itself (P2P stores into
    every peer's buffer + flags, ``ts_shift_backward_allreduce``), so the step has no separate
    collective launch at all.

    Usage (every rank, after ``init_process_group``)::

        fused = torchshifts.sharded.FusedGradWeightAllReduce()      # allocates + rendezvous
        with fused:                                                  # or fused.enable() / fused.disable()
            loss.backward()          # shift layers' weight.grad are global sums; do NOT all-reduce them again

    Every rank must run the same sequence of shift backward calls (true for replicated models).  The
    exchange buffer is symmetric memory (``torch.distributed._symmetric_memory``); with a single rank a
    plain device buffer is used and the protocol degenerates to a local copy."""

    FLAG_WORDS = 8 * 128          # (sender, CTA) flag words: 8 ranks x 128 CTAs of 32 outputs

    def __init__(self, group=None, capacity: int = 4096, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert 1 <= self.world <= 8, 'one box: at most 8 ranks'
        self.capacity = int(capacity)
        assert self.capacity <= 4096, 'at most 4096 grad_weight elements per layer'
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        floats = 2 * self.world * self.capacity + self.FLAG_WORDS
        if self.world == 1:
            self.buf = torch.zeros(floats, dtype=torch.float32, device=self.device)
            ptrs = [self.buf.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm_mem
            pg = group if group is not None else dist.group.WORLD
            with torch.cuda.device(self.device):
                self.buf = symm_mem.empty(floats, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            self.handle = symm_mem.rendezvous(self.buf, pg.group_name)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)             # every buffer is zeroed before anybody's first exchange
        self._ptrs = ptrs
        self.epoch = 0
        self._prev = None

    def peer_group(self):
        """ctypes ``ts_peer_group`` for the next backward call (bumps the epoch)."""
        from ._cabi import PeerGroup
        self.epoch += 1
        pg = PeerGroup()
        pg.world, pg.rank, pg.epoch, pg.capacity = self.world, self.rank, self.epoch, self.capacity
        flag_off = 4 * 2 * self.world * self.capacity
        for p in range(self.world):
            pg.bufs[p] = self._ptrs[p]
            pg.flags[p] = self._ptrs[p] + flag_off
        return pg

    def enable(self):
        from . import _ops
        self._prev = _ops._FUSED_ALLREDUCE
        _ops._FUSED_ALLREDUCE = self
        return self

    def disable(self):
        from . import _ops
        _ops._FUSED_ALLREDUCE = self._prev
        self._prev = None

    __enter__ = enable

    def __exit__(self, *exc):
        self.disable()
        return False
