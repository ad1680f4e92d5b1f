#!/usr/bin/env python
"""Top stall sites of an .ncu-rep (source page): tools/ncu_hot.py REP [N] [extra ncu import filters, e.g. -s 7 -c 1]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[3:], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {k: i for i, k in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print(f"total samples {tot}, instructions {len(body)}")
agg = {k: sum(int(r[col[k]] or 0) for r in body) for k in stalls}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][col["# Samples"]] or 0))[:n]:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{s:7d} {100 * s / tot:5.1f}%  #{i:5d} {r[col['Source']][:70]:70s} {top}")
