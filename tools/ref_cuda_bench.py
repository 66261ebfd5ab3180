#!/usr/bin/env python
"""Time the REFERENCE's own CUDA kernels (csrc/ops/cuda/shifts_cuda.cu, compiled unmodified for sm_100 by
`oracle/build_ref_full.sh cuda` into oracle/_ref/torchshifts_ref_cuda/_C.so) on this GPU: the "existing GPU
kernel" comparator of SURVEY.md 8c.  Runs in its OWN process and never imports this repository's torchshifts
(both register the `torchshifts::` operator namespace).

    python tools/ref_cuda_bench.py [cfg1 cfg2 cfg3 cfg3a cfg4 ...] [--json]

Prints one line per case (forward and backward separately, CUDA events, back-to-back launches after warm-up),
with the same algorithmic-bytes formula as bench.py (2e forward, 3e backward); --json prints one JSON object.
The reference has no quantized CUDA kernel (cfg5) and no bf16 (cfg2 bf16)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "oracle" / "_ref" / "torchshifts_ref_cuda" / "_C.so"

CASES = {
    "cfg1": ("cfg1 2d sparse zeros f32", (8, 64, 32, 32), 0, False),
    "cfg2": ("cfg2 1d active periodic f32", (64, 512, 4096), 2, True),
    "cfg3": ("cfg3 2d sparse zeros f32", (256, 256, 56, 56), 0, False),
    "cfg3a": ("cfg3 2d active zeros f32", (256, 256, 56, 56), 0, True),
    "cfg3r": ("cfg3 2d active reflect f32", (256, 256, 56, 56), 3, True),
    "cfg4": ("cfg4 3d active zeros f32", (32, 128, 16, 56, 56), 0, True),
    "cfg4r": ("cfg4 3d active reflect f32", (32, 128, 16, 56, 56), 3, True),
}


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    as_json = "--json" in sys.argv
    reps = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--reps=")), 10)
    want = [a for a in sys.argv[1:] if a in CASES] or ["cfg3"]
    out = {"library": str(LIB.relative_to(ROOT)), "what": "reference csrc/ops/cuda/shifts_cuda.cu, unmodified, built for sm_100",
           "cases": {}}
    if not LIB.exists():
        out["unavailable"] = "oracle/_ref/torchshifts_ref_cuda/_C.so not built (oracle/build_ref_full.sh cuda)"
        print(json.dumps(out) if as_json else out["unavailable"])
        return
    assert "torchshifts" not in sys.modules
    torch.ops.load_library(str(LIB))
    dev = torch.device("cuda:0")
    out["gpu"] = torch.cuda.get_device_name(0)
    for key in want:
        label, shape, pad, active = CASES[key]
        dim = len(shape) - 2
        torch.manual_seed(0)
        x = torch.randn(shape, device=dev)
        g = torch.randn(shape, device=dev)
        w = torch.rand(shape[1], dim, device=dev) * 2 - 1
        sp = list(shape[2:]) + [1] * (3 - dim)
        borders = torch.tensor([0, sp[0], 0, sp[1], 0, sp[2]], dtype=torch.int32, device=dev)
        fwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_forward")
        bwd = getattr(torch.ops.torchshifts, f"_shift{dim}d_backward")
        n = x.numel()
        try:
            with torch.no_grad():
                tf = timeit(lambda: fwd(x, w, borders, list(shape), pad, active), reps)
                tb = timeit(lambda: bwd(g, w, x, borders, pad, active), max(2, reps // 2), warm=1)
        except Exception as e:      # keep going: one line per case whatever happens
            out["cases"][key] = {"label": label, "error": f"{type(e).__name__}: {str(e)[:200]}"}
            if not as_json:
                print(f"{label:32s} FAILED {out['cases'][key]['error']}", flush=True)
            continue
        rec = {"label": label, "fwd_ms": tf, "bwd_ms": tb, "fwd_gbs": n * 8 / tf / 1e6, "bwd_gbs": n * 12 / tb / 1e6,
               "fwd_bwd_ms": tf + tb, "fwd_bwd_gbs": n * 20 / (tf + tb) / 1e6}
        out["cases"][key] = rec
        if not as_json:
            print(f"{label:32s} fwd {tf:8.3f} ms {rec['fwd_gbs']:7.0f} GB/s | bwd {tb:9.3f} ms {rec['bwd_gbs']:7.0f} GB/s | "
                  f"fwd+bwd {rec['fwd_bwd_gbs']:7.0f} GB/s   [reference CUDA kernels]", flush=True)
        del x, g
        torch.cuda.empty_cache()
    if as_json:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
