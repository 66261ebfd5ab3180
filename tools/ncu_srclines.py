#!/usr/bin/env python
"""Fold the stall samples of one kernel of an ncu report onto CUDA source lines (needs -lineinfo and --import-source on).
    python tools/ncu_srclines.py <rep.ncu-rep> [kernel-index (1-based)] [top N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows, cur, hdr = [], "?", None
for r in csv.reader([l for l in out.splitlines() if l.startswith('"')]):
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        rows.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), cur, int(r[0]), r[1].strip()[:120], d))
tot_s = sum(x[0] for x in rows) or 1
tot_e = sum(x[1] for x in rows) or 1
print(f"kernel {kid}: {tot_s} samples, {tot_e} warp instructions")
for s_, e_, f, ln, t, d in sorted(rows, key=lambda x: -x[0])[:top]:
    why = sorted(((int(v or 0), k.replace("stall_", "")) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
    print(f"{100 * s_ / tot_s:5.1f}% smp {100 * e_ / tot_e:5.1f}% ins  {f}:{ln}  [{', '.join(f'{k} {v}' for v, k in why)}]  {t}")
