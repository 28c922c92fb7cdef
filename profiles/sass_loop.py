#!/usr/bin/env python
"""Opcode histogram of the innermost loop that contains a given mnemonic (default FMNMX3) of one
kernel in a cubin/.so: the static evidence next to the ncu counters.
usage: sass_loop.py <file> <kernel-substring> [mnemonic]"""
import re
import subprocess
import sys


def main():
    path, kern = sys.argv[1], sys.argv[2]
    mnem = sys.argv[3] if len(sys.argv) > 3 else "FMNMX3"
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for blk in blocks:
        name = blk.split("\n", 1)[0]
        if kern not in name:
            continue
        ins = []
        for line in blk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
        # backward branches = loops; pick the smallest loop containing the mnemonic
        best = None
        for k, (addr, op, rest) in enumerate(ins):
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) < addr:
                    lo = int(t.group(1), 16)
                    body = [x for x in ins if lo <= x[0] <= addr]
                    if any(x[1].startswith(mnem) for x in body) and (best is None or len(body) < len(best)):
                        best = body
        if best is None:
            print(name, ": no loop containing", mnem)
            continue
        hist = {}
        for _, op, _ in best:
            key = op.split(".")[0]
            hist[key] = hist.get(key, 0) + 1
        print(name)
        print("  loop of %d instructions:" % len(best), sorted(hist.items(), key=lambda kv: -kv[1]))
        print("  spill traffic in loop (LDL/STL):", hist.get("LDL", 0), hist.get("STL", 0))


if __name__ == "__main__":
    main()
