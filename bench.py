#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Shift2d fwd+bwd algorithmic GB/s, N=256 C=256 56x56 fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path, host cores

One "step" = one forward + one backward of the sparse (zeros padded) Shift2d over the batch shard
this rank owns, plus -- for N>1 -- the all-reduce of the C x 2 grad_weight.  Algorithmic bytes per
input element: 2e forward + 3e backward = 20 B for fp32 (SURVEY.md 8d / BASELINE.md 2).

Keys of the JSON line (rank 0 prints exactly one line):
  value        whole-job GB/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public API with HOST (pinned) buffers: H2D of x and grad,
               forward, backward, D2H of y, grad_input and grad_weight, all inside the timed region
  roofline     dominant kernel (backward): algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline the reference CPU kernels (oracle/_ref) timed on this box's host cores, bounded sample
  clocks       nvidia-smi SM clock / throttle reasons sampled during the timed region
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "activesparseshifts-pytorch_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "shift2d_fwd_bwd_algorithmic_bandwidth"
UNIT = "GB/s"
CFG = dict(N=256, C=256, H=56, W=56)          # BASELINE.json configs[2]  (cfg3)
BYTES_PER_ELEM_FWD, BYTES_PER_ELEM_BWD = 8, 12  # fp32: 2e, 3e


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"],
                    help="weak (default): every rank owns a full BASELINE shard, N=256 images (global batch 256*G, batch-sharded); "
                         "strong: the global batch N=256 is split over the ranks")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="multi-GPU: all-reduce grad_weight with torch.distributed (NCCL) after the backward instead of the "
                         "default in-kernel exchange over NVLink peer memory (ts_shift_backward_allreduce)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample clocks during the timed region")
    ap.add_argument("--clocks", default="auto", choices=["auto", "thread", "inline", "off"],
                    help="how the SM clock / throttle reasons are sampled during the timed region: a 4 ms NVML polling "
                         "thread (default), or three NVML reads from the launching thread while the GPU works through the "
                         "queued steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-n", type=int, default=16)
    return ap.parse_args()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML
    (nvidia_ml_py; no nvidia-smi subprocess to spawn and parse).  Only rank 0 samples; its start-up
    happens BEFORE the barrier that opens the timed region, so it cannot skew the ranks."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0, period_s=0.004):
        self.rows, self.index, self.period, self.h, self.stop_flag, self.err = [], index, period_s, None, False, None

    def start(self, poll=True):
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = None
            if poll:
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
        except Exception as e:
            self.h, self.err = None, f"NVML unavailable: {e}"

    def sample_once(self):
        """one NVML read from the calling thread (used between enqueued steps)"""
        if self.h is None:
            return
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((time.perf_counter(), sm, mask))
        except Exception as e:
            self.err = str(e)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, mask))
            except Exception as e:
                self.err = str(e)
                return
            time.sleep(self.period)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "not sampled"]}
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = [r[1] for r in rows]
        reasons = sorted({name for r in rows for bit, name in self.REASONS if r[2] & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "how": "NVML (nvidia_ml_py), " + ("polled every 4 ms by a thread" if self.thread is not None else
                                                                       "read from the launching thread between enqueued steps") +
                " inside the timed region"}


# ------------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference_run(sample_n, steps, warmup, threads=None):
    """Time the reference's CPU implementation of the path on a bounded sample of the workload.

    Preference order: (1) the reference's own torch extension (oracle/_ref/torchshifts_ref/_C.so,
    its registered op + autograd, at::parallel_for over all host threads) -- only in a process that
    has NOT imported this repository's torchshifts (same op namespace); (2) the reference's
    per-element headers behind OpenMP (oracle/_ref/libref_shifts.so); (3) the oracle port."""
    import numpy as np
    cores = threads or os.cpu_count() or 1
    N, C, H, W = sample_n, CFG["C"], CFG["H"], CFG["W"]
    elems = N * C * H * W
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, C, H, W), dtype=np.float32)
    g = rng.standard_normal((N, C, H, W), dtype=np.float32)
    w = (rng.random((C, 2), dtype=np.float32) * 2 - 1)
    full = ROOT / "oracle" / "_ref" / "torchshifts_ref" / "_C.so"
    kind, how = "port", ""
    step = None
    if full.exists() and "torchshifts" not in sys.modules:
        try:
            import torch
            torch.ops.load_library(str(full))
            torch.set_num_threads(cores)
            xt, gt = torch.from_numpy(x), torch.from_numpy(g)
            wt = torch.from_numpy(w).requires_grad_(True)
            xt.requires_grad_(True)
            empty = torch.Tensor()

            def step():
                xt.grad = None; wt.grad = None
                y = torch.ops.torchshifts.shift2d(xt, wt, empty, 0, False)
                y.backward(gt)
            kind, how = "reference", "reference torch extension (unmodified csrc, op torchshifts::shift2d + its autograd), at::parallel_for"
        except Exception as e:  # fall through to the header build
            step, how = None, f"(extension unusable: {e}) "
    if step is None:
        from oracle.oracle import Oracle
        if Oracle.available("reference"):
            orc = Oracle("reference", threads=cores)
            kind, how = "reference", how + "reference per-element headers (kernels/shifts_kernels.h) behind OpenMP, oracle/_ref/libref_shifts.so"
        else:
            orc = Oracle("port", threads=1)
            cores = 1
            kind, how = "port", how + "oracle/shift_oracle.c (scalar C restatement)"

        def step():
            orc.forward(x, w, 0, False)
            orc.backward(g, x, w, 0, False)
    for _ in range(max(1, warmup)):
        step()
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    gbs = elems * (BYTES_PER_ELEM_FWD + BYTES_PER_ELEM_BWD) / t / 1e9
    sample = f"Shift2d SSL zeros fwd+bwd on N={N} of 256 images (C=256, 56x56, fp32), median of {len(times)} runs; {how}"
    return {"value": gbs, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "ms_per_step": t * 1e3,
            "elements_per_s": elems / t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: at most ~25 steps of a 16-image sample, whatever --steps says (a few minutes at worst)
    steps, warmup = min(args.steps, 25), min(max(args.warmup, 1), 3)
    r = cpu_reference_run(args.cpu_sample_n, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3 Shift2d SSL zeros fwd+bwd N=256 C=256 56x56 fp32 (bounded sample, see cpu_baseline.sample)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "elements_per_s": r["elements_per_s"],
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torchshifts  # noqa: F401
    from torchshifts.extension import native
    from torchshifts.functional import shift2d_func

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = native().lib

    N = CFG["N"] if args.scaling == "weak" else CFG["N"] // world
    C, H, W = CFG["C"], CFG["H"], CFG["W"]
    elems_rank = N * C * H * W
    elems_job = elems_rank * world
    torch.manual_seed(rank)
    x = torch.randn(N, C, H, W, device=dev)
    g = torch.randn(N, C, H, W, device=dev)
    torch.manual_seed(0)
    w = (torch.rand(C, 2, device=dev) * 2 - 1).requires_grad_(True)   # module default init U(-1,1), same on every rank
    xr = x.requires_grad_(True)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    bwd_pairs = []
    fused, fused_note = None, ""
    if world > 1 and not args.nccl_allreduce:
        # every rank must take the same branch: agree on the outcome of the symmetric-memory set-up
        try:
            from torchshifts.sharded import FusedGradWeightAllReduce
            fused = FusedGradWeightAllReduce(capacity=4096, device=dev)
            ok = torch.ones(1, device=dev)
        except Exception as e:      # no symmetric memory on this box: NCCL all-reduce instead
            fused, fused_note = None, f" (in-kernel exchange unavailable: {type(e).__name__}: {str(e)[:120]})"
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) < 1:
            fused = None
        else:
            fused.enable()

    def step(timed=False):
        xr.grad = None; w.grad = None
        y = shift2d_func(xr, w, 0, False)
        if timed:
            a, b = ev(), ev(); a.record()
        y.backward(g)
        if timed:
            b.record(); bwd_pairs.append((a, b))
        if world > 1 and fused is None:
            dist.all_reduce(w.grad)          # the one collective of the path: C x 2 floats
        return y                             # (fused: w.grad is already the global sum)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    mode = "off" if args.no_clocks else args.clocks
    if mode == "auto":
        mode = "thread"
    if rank == 0 and mode != "off":
        sampler.start(poll=(mode == "thread")); time.sleep(0.15)
    inline_at = {args.steps // 3, (2 * args.steps) // 3, args.steps - 1} if (rank == 0 and mode == "inline") else set()
    if world > 1:
        dist.barrier()      # AFTER the sampler start-up (rank 0 only): every rank enters the timed region together
    launches0 = lib.ts_launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    start, end = ev(), ev()
    start.record()
    for i in range(args.steps):
        step(timed=True)
        if i in inline_at:
            sampler.sample_once()        # the GPU is still working through the queued steps
    end.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    launches = lib.ts_launch_count() - launches0
    ms = start.elapsed_time(end) / args.steps
    bwd_ms = statistics.mean(a.elapsed_time(b) for a, b in bwd_pairs)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    path = lib.ts_last_kernel_path()
    if world > 1:
        t = torch.tensor([ms, bwd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, bwd_ms = t.tolist()
    value = elems_job * (BYTES_PER_ELEM_FWD + BYTES_PER_ELEM_BWD) / (ms * 1e-3) / 1e9

    # ---- e2e: public API, host (pinned) buffers, copies inside the timed region -----------------
    e2e = None
    if fused is not None:
        fused.disable()       # the host pipeline reduces its per-chunk grad_weight once per step (NCCL below)
    if not args.no_e2e:
        from torchshifts.host import HostShift2dPipeline
        pipe = HostShift2dPipeline(N, C, H, W, device=dev)
        pipe.x_host.copy_(x.detach().cpu()); pipe.g_host.copy_(g.cpu())
        wh = w.detach().clone()
        for _ in range(2):
            pipe.forward_backward(wh, 0, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            gw = pipe.forward_backward(wh, 0, False)
            if world > 1:
                dist.all_reduce(gw)
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) / k * 1e3
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": elems_job * 20 / (e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
               "how": pipe.describe()}
        del pipe

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    bwd_bytes = elems_rank * BYTES_PER_ELEM_BWD
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("backward_dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"cfg3 Shift2d SSL zeros fwd+bwd, global N={N * world} C={C} {H}x{W} fp32, batch-sharded over {world} GPU(s)",
                   "per_gpu_batch": N, "padding": "zeros", "active": False, "weights": "U(-1,1)",
                   "l2": "inputs (822 MB per tensor at N=256) are larger than the 126 MB L2; no flush needed",
                   "kernel_path": {1: "generic", 2: "staged (cp.async.bulk + mbarrier)",
                                   3: "TMA tensor boxes (cp.async.bulk.tensor.5d, shift + zero pad by the copy engine)"}.get(path, str(path)),
                   "collective": ("none" if world == 1 else "torch.distributed all_reduce (NCCL) of grad_weight [C,2]" + fused_note if fused is None else
                                  "grad_weight [C,2] summed over the ranks inside the pass-2 reduction kernel (P2P stores + flags "
                                  "over NVLink peer memory, ts_shift_backward_allreduce); no separate collective launch")},
        "elements_per_s": elems_job / (ms * 1e-3),
        "frac_of_hbm_peak": value / world / peak,
        "roofline": {"bound": "hbm", "kernel": "shift backward (grad_input + grad_weight partials) + pass-2 reduce",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_launch": bwd_bytes, "avg_launch_ms": bwd_ms, "traffic": traffic},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "1",
                                  "--cpu-sample-n", str(args.cpu_sample_n)], capture_output=True, text=True, timeout=600)
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
