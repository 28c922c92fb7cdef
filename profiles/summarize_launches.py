#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,...] --csv` launch list of profiles/prof_moe.py by (kernel, grid):
python profiles/summarize_launches.py launches.csv  ->  one line per group, share of the forward pass (after the last
pack_mups_bf16_kernel launch)."""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rows = collections.OrderedDict()
for x in csv.DictReader(lines):
    r = rows.setdefault(x["ID"], {"name": x["Kernel Name"].split("(")[0].replace("mups::", "").replace("void ", ""), "grid": x["Grid Size"]})
    r[x["Metric Name"]] = float(x["Metric Value"])
rows = list(rows.values())
start = max(i for i, r in enumerate(rows) if "pack_mups" in r["name"])
fw = rows[start:]
tot = sum(r["gpu__time_duration.sum"] for r in fw)
print("launches %d, total %.2f ms" % (len(fw), tot / 1e6))
agg = collections.OrderedDict()
for r in fw:
    a = agg.setdefault((r["name"][:34], r["grid"]), [0, 0.0, 0.0])
    a[0] += 1
    a[1] += r["gpu__time_duration.sum"]
    a[2] += r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * r["gpu__time_duration.sum"]
for k, (c, t, ta) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%-36s %-16s x%-3d %8.3f ms %5.1f%%  avg %7.1f us  tensor active %4.1f%%" % (k[0], k[1], c, t / 1e6, 100 * t / tot, t / c / 1e3, ta / t))
