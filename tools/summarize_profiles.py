#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the small, tracked summaries under profiles/:
    python tools/summarize_profiles.py <round-tag> <launches.csv> <prof.ncu-rep>
  profiles/<tag>_launches.md      per-kernel share of a bench.py step (ncu gpu__time_duration.sum list)
  profiles/<tag>_kernels.csv      the ncu --set full metrics that the roofline numbers come from
  profiles/traffic.json           dram bytes per launch of the dominant kernels (bench.py reads it)"""
import csv
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag, launches, rep = sys.argv[1], Path(sys.argv[2]), Path(sys.argv[3])
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

# ---- launch list ---------------------------------------------------------------------------
rows = []
with open(launches) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000
        rows.append((int(r["ID"]), r["Kernel Name"], us, r["Grid Size"], r["Block Size"]))
short = lambda k: ("k_" + k.split("k_", 1)[1].split("(")[0]) if "k_" in k else k.split("<")[0].split("(")[0][-60:]
ours = [(i, short(k), us) for i, k, us, *_ in rows if "k_tma" in k or "k_staged" in k or "k_reduce" in k or "generic" in k]
agg = defaultdict(lambda: [0, 0.0])
for _, k, us in ours:
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
with open(out / f"{tag}_launches.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (cold-cache, serialised replays: compare SHARES).\n")
    f.write(f"{len(rows)} launches captured, {len(ours)} from libtorchshifts_b200.so (the rest: torch RNG / fill kernels that create the synthetic inputs).\n\n")
    f.write("| kernel | launches | total us | mean us | share of our kernels |\n|---|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / tot:.1f} % |\n")
    f.write("\nPer-launch list (our kernels, in launch order):\n\n| id | kernel | us |\n|---|---|---|\n")
    for i, k, us in ours:
        f.write(f"| {i} | `{k}` | {us:.1f} |\n")

# ---- full-set metrics ------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
hdr, units = r[0], r[1]
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
traffic = {}
with open(out / f"{tag}_kernels.csv", "w", newline="") as f:
    wtr = csv.writer(f)
    wtr.writerow(["metric", "unit"] + [short(row[hdr.index("Kernel Name")]) for row in r[2:]])
    for k in keys[1:]:
        if k in hdr:
            wtr.writerow([k, units[hdr.index(k)]] + [row[hdr.index(k)] for row in r[2:]])
to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in r[2:]:
    name = short(row[hdr.index("Kernel Name")])
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(row[hdr.index(k)].replace(",", "")) * to_bytes[units[hdr.index(k)]]
    traffic[name] = b
tj = {"source": f"profiles/{tag}_kernels.csv (ncu --set full --clock-control none, python tools/profile_step.py cfg3)",
      "dram_bytes_per_launch": traffic}
for name, b in traffic.items():
    if "backward" in name:
        tj["backward_dram_bytes_per_launch"] = b
    if "gather" in name:
        tj["forward_dram_bytes_per_launch"] = b
(out / "traffic.json").write_text(json.dumps(tj, indent=1) + "\n")
print(open(out / f"{tag}_launches.md").read()[:1500])
print(open(out / f"{tag}_kernels.csv").read())
