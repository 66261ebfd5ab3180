#!/usr/bin/env python
"""Compact summary of an ncu report (run on the GPU box right after the capture; the .ncu-rep files are too large to
bring back): per kernel the metrics the roofline discussion uses, the warp-stall breakdown, the executed-opcode mix and
the hottest SASS lines.     python tools/ncu_extract.py <rep.ncu-rep> <out.md> [title]"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
lines = [f"# {title}\n", "`ncu --set full --clock-control none` (cold-cache, serialised replays).\n"]
for i, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    lines.append(f"\n## launch {i}: `{d.get('Kernel Name', '?')}`\n\n| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in d and d[w] != "":
            lines.append(f"| {w} | {d[w]} | {units[hdr.index(w)]} |")
    stalls = [(float(d[k] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio")]
    lines.append("\nwarps stalled per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{i + 1}"], capture_output=True, text=True).stdout
    srows = list(csv.reader([l for l in src.splitlines() if l.startswith('"')]))
    if len(srows) > 2:
        sh = srows[1] if "Source" in srows[1] else srows[0]
        try:
            isrc, ismp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
        except ValueError:
            continue
        data, seen = [], set()
        for rr in srows[2:]:
            if len(rr) <= iex or rr[0] in seen:
                continue
            seen.add(rr[0])
            try:
                data.append((int(rr[ismp] or 0), int(rr[iex] or 0), rr[isrc].strip()))
            except ValueError:
                pass
        tot_s, tot_e = sum(x[0] for x in data) or 1, sum(x[1] for x in data) or 1
        ops = collections.Counter()
        for s_, e_, t_ in data:
            parts = t_.split()
            op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
            ops[op.split(".")[0]] += e_
        lines.append(f"\nexecuted warp instructions by opcode ({tot_e} total): " + ", ".join(f"{k} {100 * v / tot_e:.1f}%" for k, v in ops.most_common(14)))
        lines.append("\nhottest SASS lines (share of stall samples):\n")
        for s_, e_, t_ in sorted(data, reverse=True)[:10]:
            lines.append(f"* {100 * s_ / tot_s:.1f}% `{t_}` (executed {e_})")
    # the same samples folded onto CUDA source lines (needs -lineinfo and --import-source on)
    cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-id", f":::{i + 1}"],
                        capture_output=True, text=True).stdout
    crows = list(csv.reader([l for l in cs.splitlines() if l.startswith('"')]))
    if len(crows) > 2:
        ch = next((r for r in crows[:3] if "Source" in r and "# Samples" in r), None)
        if ch:
            isrc, ismp, iex = ch.index("Source"), ch.index("# Samples"), ch.index("Instructions Executed")
            cdata = []
            for rr in crows:
                if rr is ch or len(rr) <= iex:
                    continue
                try:
                    cdata.append((int(rr[ismp] or 0), int(rr[iex] or 0), rr[0], rr[isrc].strip()[:110]))
                except ValueError:
                    pass
            tot = sum(x[0] for x in cdata) or 1
            lines.append("\nhottest CUDA source lines (share of stall samples, warp instructions executed):\n")
            for s_, e_, ln, t_ in sorted(cdata, reverse=True)[:16]:
                lines.append(f"* {100 * s_ / tot:.1f}% line {ln}: `{t_}` ({e_})")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
