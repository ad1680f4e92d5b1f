#!/usr/bin/env python
"""Where a kernel wastes shared-memory wavefronts / global sectors, from the SOURCE page of an .ncu-rep (read offline,
no GPU): per kernel the totals of `L1 Wavefronts Shared (Excessive)` and `L2 Theoretical Sectors Global (Excessive)`,
the instructions that cause them, and the executed-instruction mix by opcode (spin loops show up as BRA / SYNCS / YIELD).
This is how the 2-way bank conflicts of the fused PointConv's operand stores and the misaligned LDGSTS staging rows of the
cost volume were found.      tools/ncu_excess.py REP [top N]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 6
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for n, hi in enumerate(heads):
    end = heads[n + 1] if n + 1 < len(heads) else len(rows)
    h = rows[hi]
    col = {k: i for i, k in enumerate(h)}
    body = [r for r in rows[hi + 1:end] if len(r) == len(h)]

    def num(r, k):
        try:
            return float(r[col[k]] or 0)
        except (ValueError, KeyError):
            return 0.0

    print(f"== kernel {n}: {len(body)} instructions, {sum(num(r, 'Instructions Executed') for r in body):.0f} executed (warp level)")
    for total, excess in (("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"),
                          ("L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Excessive")):
        t, x = sum(num(r, total) for r in body), sum(num(r, excess) for r in body)
        print(f"   {total}: {t:.0f}, excessive {x:.0f} ({100 * x / max(t, 1):.1f} %)")
        for i, r in sorted(enumerate(body), key=lambda ir: -num(ir[1], excess))[:top]:
            if num(r, excess) > 0:
                print(f"      #{i:5d} {r[col['Source']][:64]:64s} executed {num(r, 'Instructions Executed'):9.0f}  "
                      f"total {num(r, total):10.0f}  excessive {num(r, excess):10.0f}")
    mix = Counter()
    for r in body:
        src = r[col["Source"]].split()
        if src:
            op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
            mix[op.split(".")[0]] += num(r, "Instructions Executed")
    tot = sum(mix.values()) or 1.0
    print("   instruction mix: " + "  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in mix.most_common(10)))
