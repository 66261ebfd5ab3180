#!/usr/bin/env bash
# Builds the UNMODIFIED reference torch extension (CPU kernels + autograd + quantized) from
# /root/reference into oracle/_ref/torchshifts_ref/_C.so.  TEST INFRASTRUCTURE ONLY.
#
# The reference sources are never copied into this repository: the build happens in a scratch
# copy under /tmp (the reference tree is read-only and setup.py writes version.py / build/ into
# its own tree) and only the resulting shared object is brought back (git-ignored, but it travels
# to the GPU box with gpurun).  One compat patch is applied to the scratch copy:
# csrc/ops/quantized/shifts_quantized.cpp:126 hands a std::string to AT_DISPATCH_QINT_TYPES,
# which torch >= 2.x only accepts as a const char*.
set -euo pipefail
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${TS_REFERENCE_DIR:-/root/reference}"
# `build_ref_full.sh cuda` builds the reference WITH its own CUDA kernels for sm_100 (csrc/ops/cuda/shifts_cuda.cu,
# FORCE_CUDA=1 TORCH_CUDA_ARCH_LIST=10.0): the "existing GPU kernel" comparator that tools/ref_cuda_bench.py
# times on the B200 box in a separate process.  It is never loaded next to the product library.
VARIANT="${1:-cpu}"
if [ "$VARIANT" = "cuda" ]; then
    OUT="$REPO/oracle/_ref/torchshifts_ref_cuda"; export FORCE_CUDA=1 TORCH_CUDA_ARCH_LIST="10.0"
else
    OUT="$REPO/oracle/_ref/torchshifts_ref"; export FORCE_CUDA=0
fi
[ -d "$SRC/torchshifts/csrc" ] || { echo "reference tree not found at $SRC" >&2; exit 3; }
WORK="$(mktemp -d /tmp/tsref.XXXXXX)"
trap 'rm -rf "$WORK"' EXIT
cp -r "$SRC/." "$WORK/"
chmod -R u+w "$WORK"
cd "$WORK"
sed -i 's/AT_DISPATCH_QINT_TYPES(input.scalar_type(), name,/AT_DISPATCH_QINT_TYPES(input.scalar_type(), "q_shiftnd_cpu",/' \
    torchshifts/csrc/ops/quantized/shifts_quantized.cpp
python setup.py build_ext --inplace > "$WORK/build.log" 2>&1 || { tail -50 "$WORK/build.log" >&2; exit 4; }
mkdir -p "$OUT"
cp torchshifts/_C*.so "$OUT/_C.so"
echo "built $OUT/_C.so"
