#!/usr/bin/env python
"""Per-source-line share of the executed warp instructions and of the stall samples of one kernel in an .ncu-rep
(captured with --set full --import-source on, library built with -lineinfo): joins the SASS page of the report with the
line table nvdisasm prints for the same function of the built library.

usage: ncu_lines.py <rep> <kernel-regex> <mangled-function-substring> <cubin-substring> <source-file> [top N]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kregex, mangled, cubin_key, srcfile = sys.argv[1:6]
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kregex], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci, cs, ca = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "nesti-net_b200", "libmups_b200.so")], cwd=d, capture_output=True)
        cub = next(os.path.join(d, f) for f in sorted(os.listdir(d)) if cubin_key in f and f.count("-") == 0)
        sass = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(sass) if ".section" in l and mangled in l and ".text." in l)
    off2line, cur = {}, None
    pl, pi = re.compile(r'//## File "([^"]+)", line (\d+)'), re.compile(r"/\*([0-9a-f]{4,})\*/\s+\S")
    for l in sass[start + 1:]:
        if ".section" in l:
            break
        m = pl.search(l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = pi.search(l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    body = [r for r in rows[h + 1:] if len(r) > ci and r[ca].startswith("0x")]
    base = int(body[0][ca], 16)
    agg, tot, ts = {}, 0.0, 0.0
    for r in body:
        ln = off2line.get(int(r[ca], 16) - base)
        inst, sm = float(r[ci] or 0), float(r[cs] or 0)
        tot += inst
        ts += sm
        a = agg.setdefault(ln, [0.0, 0.0])
        a[0] += inst
        a[1] += sm
    src = open(srcfile).read().split("\n")
    print("kernel %s: %.4g warp instructions, %d stall samples\n" % (kregex, tot, ts))
    print("| line | warp instructions | stall samples | source |\n|---|---|---|---|")
    for ln, (a, b) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        t = src[ln[1] - 1].strip()[:110] if ln and ln[0] == os.path.basename(srcfile) else str(ln)
        print("| %s | %.1f %% | %.1f %% | `%s` |" % (ln[1] if ln else "?", 100 * a / tot, 100 * b / max(ts, 1), t.replace("|", "\\|")))


if __name__ == "__main__":
    main()
