#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (captured with --set full --import-source on) as markdown:
key counters (duration, DRAM bytes, pipe utilisation, stall reasons, occupancy limits) plus the
instruction/stall share of the code regions between barriers.
usage: summarize_ncu.py <rep> <title> [queries-per-launch] [kernel-name-regex] > profiles/rNN_<kernel>.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


KERNEL = None


def page(rep, name):
    cmd = ["ncu", "-i", rep, "--page", name, "--csv"] + (["-k", "regex:" + KERNEL] if KERNEL else [])
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    nq = float(sys.argv[3]) if len(sys.argv) > 3 else None
    global KERNEL
    KERNEL = sys.argv[4] if len(sys.argv) > 4 else None
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("# %s\n" % title)
    print("Source: `%s` (ncu --set full --clock-control none --import-source on; cold-cache, serialised replay).\n" % rep.split("/")[-1])
    print("Kernel: `%s`\n" % m.get("Kernel Name", ("?", ""))[0])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in m:
            print("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
    def num(k):
        v, u = m[k]
        f = float(v.replace(",", ""))
        return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    if nq and "dram__bytes_read.sum" in m:
        rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
        print("\nDRAM traffic per query point: %.0f B read + %.0f B written = %.0f B (%g query points in this launch)."
              % (rd / nq, wr / nq, (rd + wr) / nq, nq))
        print("<!-- dram_bytes_per_query=%.1f -->" % ((rd + wr) / nq))
    stalls = sorted(((float(v[0]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, v in m.items()
                     if "issue_stalled" in h and "per_issue_active.ratio" in h), reverse=True)
    print("\nStall reasons (warps stalled per issue-active cycle): " +
          ", ".join("%s %.2f" % (n, x) for x, n in stalls if x >= 0.05) + "\n")
    src = page(rep, "source")
    if len(src) > 3:
        h = src[1]
        data = src[2:]
        isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        tot = sum(int(r[iex]) for r in data) or 1
        tots = sum(int(r[ism]) for r in data) or 1
        print("Regions between barriers (SASS rows, share of executed warp instructions, share of stall samples):\n")
        print("| SASS rows | instructions | samples | ends with |\n|---|---|---|---|")
        cum = cs = 0
        start = 0
        for k, r in enumerate(data):
            cum += int(r[iex]); cs += int(r[ism])
            if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc] or k == len(data) - 1:
                if cum > 0.004 * tot or cs > 0.004 * tots:
                    print("| %d-%d | %.1f %% | %.1f %% | `%s` |" % (start, k, 100.0 * cum / tot, 100.0 * cs / tots, r[isrc].strip()[:40]))
                start, cum, cs = k + 1, 0, 0
        ops = {}
        for r in data:
            op = r[isrc].strip().split(" ")
            op = [x for x in op if x and not x.startswith("@")]
            if op:
                key = op[0].split(".")[0]
                ops[key] = ops.get(key, 0) + int(r[iex])
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
        print("\nExecuted warp instructions by opcode: " + ", ".join("%s %.1f %%" % (k, 100.0 * v / tot) for k, v in top))


if __name__ == "__main__":
    main()
