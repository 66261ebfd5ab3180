"""Host-resident tensors through the GPU path: ``HostShift2dPipeline`` / ``HostShiftPipeline``.

The reference's CPU operator takes host tensors and returns host tensors.  This build has no CPU
compute path, so the equivalent service for host-resident data is a streamed offload: the batch
is cut into chunks, and for every chunk

    copy-in stream :  x[chunk], grad[chunk]  pinned host -> device       (H2D copy engine)
    compute stream :  forward + backward kernels (the registered torchshifts ops)
    copy-out stream:  y[chunk], grad_input[chunk]  device -> pinned host (D2H copy engine)

run concurrently on three CUDA streams with a small ring of device buffers, so the two PCIe
directions and the kernels overlap; grad_weight is accumulated on the device across chunks in a
fixed order (deterministic) and returned as a device tensor.  bench.py's ``e2e`` number is this
call, with every host<->device byte inside the timed region.
"""
import torch

from .functional import shift1d_func, shift2d_func, shift3d_func  # noqa: F401  (public API re-export)

_OPS = {1: ('_shift1d_forward', '_shift1d_backward'), 2: ('_shift2d_forward', '_shift2d_backward'),
        3: ('_shift3d_forward', '_shift3d_backward')}


class HostShiftPipeline:
    def __init__(self, N, C, spatial, device, dtype=torch.float32, chunk=16, slots=4):
        self.N, self.C, self.spatial = int(N), int(C), tuple(int(s) for s in spatial)
        self.dim = len(self.spatial)
        assert self.dim in (1, 2, 3)
        self.device = torch.device(device)
        self.dtype = dtype
        self.chunk = max(1, min(int(chunk), self.N))
        self.slots = slots
        shape = (self.N, self.C) + self.spatial
        pin = dict(dtype=dtype, pin_memory=True)
        self.x_host = torch.empty(shape, **pin)
        self.g_host = torch.empty(shape, **pin)
        self.y_host = torch.empty(shape, **pin)
        self.gi_host = torch.empty(shape, **pin)
        self.gw_host = torch.empty((self.C, self.dim), **pin)
        self.gw_device = None
        cshape = (self.chunk, self.C) + self.spatial
        self._xd = [torch.empty(cshape, dtype=dtype, device=self.device) for _ in range(slots)]
        self._gd = [torch.empty(cshape, dtype=dtype, device=self.device) for _ in range(slots)]
        # distinct priorities: streams of different priority never share a hardware queue, so the two copy directions and the
        # kernels cannot serialise behind each other through a false queue dependency (see bench.py)
        self._s_in = torch.cuda.Stream(self.device, priority=0)
        self._s_comp = torch.cuda.Stream(self.device, priority=-1)
        self._s_out = torch.cuda.Stream(self.device, priority=-2)
        self._borders = torch.tensor([0, self.spatial[0], 0, self.spatial[1] if self.dim > 1 else 1,
                                      0, self.spatial[2] if self.dim > 2 else 1], dtype=torch.int32)
        esz = torch.empty((), dtype=dtype).element_size()
        numel = self.x_host.numel()
        self.h2d_bytes = 2 * numel * esz
        self.d2h_bytes = 2 * numel * esz + self.C * self.dim * esz

    def describe(self):
        return (f"HostShiftPipeline: pinned host x/grad -> device in chunks of {self.chunk} images on a copy-in stream, "
                f"torchshifts::_shift{self.dim}d_forward/_backward on a compute stream, y/grad_input -> pinned host on a "
                f"copy-out stream ({self.slots}-slot ring); grad_weight summed on device and copied to pinned host "
                f"(read_back_grad_weight)")

    @torch.no_grad()
    def forward_backward(self, weight, padding_mode=0, active_flag=False):
        """Runs y = shift(x_host), (grad_input, grad_weight) = backward(g_host); fills ``y_host`` and
        ``gi_host`` (complete once the current stream is synchronised) and returns grad_weight [C, dim]
        as a device tensor ordered on the current stream."""
        fwd = getattr(torch.ops.torchshifts, _OPS[self.dim][0])
        bwd = getattr(torch.ops.torchshifts, _OPS[self.dim][1])
        cur = torch.cuda.current_stream(self.device)
        for s in (self._s_in, self._s_comp, self._s_out):
            s.wait_stream(cur)
        gw_total = torch.zeros(self.C, self.dim, dtype=self.dtype, device=self.device)
        self.gw_device = gw_total
        self._s_comp.wait_stream(cur)
        free = [None] * self.slots           # event: slot's device inputs may be overwritten
        # Outputs of a chunk (allocated by the operators on the compute stream) are kept alive per slot until the copy-out of
        # that slot has finished AND the compute stream has been ordered behind it; only then are the references dropped, so
        # the caching allocator hands the same few blocks back to the compute stream.  (tensor.record_stream() instead made
        # the allocator decide by event queries at allocation time: whenever a copy-out was still in flight it called
        # cudaMalloc -- a device-wide synchronisation -- and one pipeline instance in four ran at half speed or worse.)
        held = [None] * self.slots           # (y, grad_input, copy-out-done event) of the slot's previous chunk
        nchunks = (self.N + self.chunk - 1) // self.chunk
        for i in range(nchunks):
            lo, hi = i * self.chunk, min(self.N, (i + 1) * self.chunk)
            k = i % self.slots
            n = hi - lo
            with torch.cuda.stream(self._s_in):
                if free[k] is not None:
                    self._s_in.wait_event(free[k])
                xd, gd = self._xd[k][:n], self._gd[k][:n]
                xd.copy_(self.x_host[lo:hi], non_blocking=True)
                gd.copy_(self.g_host[lo:hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self._s_in)
            with torch.cuda.stream(self._s_comp):
                self._s_comp.wait_event(ready)
                if held[k] is not None:
                    self._s_comp.wait_event(held[k][2])      # the slot's previous outputs have left the device ...
                    held[k] = None                           # ... so their blocks may serve this chunk's outputs
                new_size = [n, self.C] + list(self.spatial)
                y = fwd(xd, weight, self._borders, new_size, padding_mode, active_flag)
                gi, gw = bwd(gd, weight, xd, self._borders, padding_mode, active_flag)
                gw_total += gw
                done = torch.cuda.Event()
                done.record(self._s_comp)
                free[k] = done
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(done)
                self.y_host[lo:hi].copy_(y, non_blocking=True)
                self.gi_host[lo:hi].copy_(gi, non_blocking=True)
                out_done = torch.cuda.Event()
                out_done.record(self._s_out)
                held[k] = (y, gi, out_done)
        for k in range(self.slots):                          # order the compute stream behind the last copies before the references go
            if held[k] is not None:
                self._s_comp.wait_event(held[k][2])
        held = None
        cur.wait_stream(self._s_comp)
        cur.wait_stream(self._s_out)
        cur.wait_stream(self._s_in)
        return gw_total

    def read_back_grad_weight(self, gw=None):
        """grad_weight [C, dim] -> pinned host (``gw_host``), ordered on the current stream; this is the
        ``C*dim*esize`` bytes ``d2h_bytes`` counts.  ``gw``: the tensor to read back (default: the device sum of
        the last ``forward_backward``; multi-GPU callers pass the all-reduced one)."""
        self.gw_host.copy_(self.gw_device if gw is None else gw, non_blocking=True)
        return self.gw_host


class HostShift2dPipeline(HostShiftPipeline):
    def __init__(self, N, C, H, W, device, dtype=torch.float32, chunk=16, slots=4):
        super().__init__(N, C, (H, W), device, dtype, chunk, slots)


class GraphedShiftStep:
    """Forward + backward of one shift layer captured ONCE into CUDA graphs and replayed: the whole step costs
    one (``split=False``) or two (``split=True``: forward graph, backward graph -- so the two can be timed
    separately) graph launches instead of ~270 us of Python / dispatcher / launch work per step, which is what
    bounds small per-GPU batches (cfg1; cfg3 split over 8 GPUs = 32 images per GPU).

    The C ABI only enqueues kernels on the caller's stream, tensor maps are encoded on the host and passed by
    value, and the in-kernel grad_weight exchange keeps its call counter on the device, so the capture includes
    the fused all-reduce when a :class:`torchshifts.sharded.FusedGradWeightAllReduce` is enabled around the
    construction AND the replays (every rank must then construct and replay in lock-step).

    ``x``, ``weight`` and ``grad_out`` are the static buffers: overwrite them in place (``copy_``) between
    replays; ``y``, ``grad_input`` and ``grad_weight`` are static outputs, valid after ``replay()`` in stream
    order."""

    def __init__(self, x, weight, grad_out, padding_mode=0, active_flag=False, borders=None, split=False, warmup=3):
        dim = x.dim() - 2
        func = {1: shift1d_func, 2: shift2d_func, 3: shift3d_func}[dim]
        self.x = x.detach().requires_grad_(True)
        self.weight = weight.detach().requires_grad_(True)
        self.grad_out = grad_out
        dev = x.device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):       # also settles allocator state and lazy initialisation
                y = func(self.x, self.weight, padding_mode, active_flag, borders)
                torch.autograd.grad(y, (self.x, self.weight), grad_out)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.forward_graph = torch.cuda.CUDAGraph()
        self.backward_graph = None
        if split:
            with torch.cuda.graph(self.forward_graph):
                self.y = func(self.x, self.weight, padding_mode, active_flag, borders)
            self.backward_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.backward_graph, pool=self.forward_graph.pool()):
                self.grad_input, self.grad_weight = torch.autograd.grad(self.y, (self.x, self.weight), grad_out)
        else:
            with torch.cuda.graph(self.forward_graph):
                self.y = func(self.x, self.weight, padding_mode, active_flag, borders)
                self.grad_input, self.grad_weight = torch.autograd.grad(self.y, (self.x, self.weight), grad_out)

    def replay_forward(self):
        self.forward_graph.replay()

    def replay_backward(self):
        if self.backward_graph is not None:
            self.backward_graph.replay()

    def replay(self):
        self.forward_graph.replay()
        if self.backward_graph is not None:
            self.backward_graph.replay()
        return self.y, self.grad_input, self.grad_weight
