#!/usr/bin/env python
"""SASS opcode histogram of libtorchshifts_b200.so (cuobjdump -sass): the mnemonics that prove the data movers are
Blackwell-native (UTMALDG = TMA tensor load, UBLKCP = bulk async copy, SYNCS = mbarrier, STG.E.EF.128 = 128-bit
streaming store), per kernel family.     python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "activesparseshifts-pytorch_b200" / "torchshifts" / "libtorchshifts_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
WATCH = ["UTMALDG", "UBLKCP", "SYNCS", "STG.E.EF.128", "STG.E.EF.64", "LDS.128", "LDGSTS", "BAR.SYNC", "SHFL", "DADD", "LDL", "STL", "FFMA", "FMUL", "FADD"]
fam = lambda fn: next((k for k in ("k_tma", "k_staged", "k_halo", "k_flat", "k_gather_nhwc", "k_nhwc_to_nchw", "k_reduce", "generic") if k in fn), "other")
per = defaultdict(Counter)
funcs = Counter()
cur = "?"
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = fam(m.group(1))
        funcs[cur] += 1
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    per[cur]["total"] += 1
    for w in WATCH:
        if op.startswith(w):
            per[cur][w] += 1
print(f"# SASS opcode histogram of `{LIB.relative_to(ROOT)}` (sm_100a)\n")
print("`cuobjdump -sass` static instruction counts per kernel family (all template instantiations).\n")
print("| family | kernels | instructions | " + " | ".join(WATCH) + " |")
print("|---|---|---|" + "---|" * len(WATCH))
for k in sorted(per, key=lambda k: -per[k]["total"]):
    print(f"| {k} | {funcs[k]} | {per[k]['total']} | " + " | ".join(str(per[k][w]) for w in WATCH) + " |")
tot = Counter()
for k in per:
    tot.update(per[k])
print(f"| **all** | {sum(funcs.values())} | {tot['total']} | " + " | ".join(str(tot[w]) for w in WATCH) + " |")
