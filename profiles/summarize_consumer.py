#!/usr/bin/env python
"""Markdown table of the consumer's kernels from `ncu -i <rep> --page raw --csv` (stdin or file argument):
ncu -i r02_conv3.ncu-rep --page raw --csv | python profiles/summarize_consumer.py"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
to_us = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[units[idx["gpu__time_duration.sum"]]]
to_gb = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
xbar_gb = to_gb[units[idx["l1tex__m_xbar2l1tex_read_bytes.sum"]]]
dr_mb = to_gb[units[idx["dram__bytes_read.sum"]]] * 1e3
dw_mb = to_gb[units[idx["dram__bytes_write.sum"]]] * 1e3
print("| # | kernel | grid | smem KB | duration | tensor pipe active | L2 sectors (hit rate) | L2 -> SM | DRAM read / written | issue active |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(rows[2:]):
    g = lambda k: r[idx[k]]  # noqa: E731
    t = float(g("gpu__time_duration.sum")) * to_us
    xbar = float(g("l1tex__m_xbar2l1tex_read_bytes.sum")) * xbar_gb
    print("| %d | `%s` | %s | %.0f | %.1f us | %.1f %% | %.1f M (%.0f %%) | %.2f GB = %.1f TB/s | %.0f / %.0f MB | %.1f %% |" % (
        n, g("Kernel Name").split("(")[0].replace("mups::", "").replace("void ", ""), g("launch__grid_size"),
        float(g("launch__shared_mem_per_block_dynamic")), t, float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
        float(g("lts__t_sectors.sum")) / 1e6, min(100.0, float(g("lts__t_sector_hit_rate.pct"))), xbar, xbar / t * 1e3,
        float(g("dram__bytes_read.sum")) * dr_mb, float(g("dram__bytes_write.sum")) * dw_mb, float(g("smsp__issue_active.avg.pct_of_peak_sustained_active"))))
