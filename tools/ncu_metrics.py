#!/usr/bin/env python
"""Print selected raw metrics of an .ncu-rep: tools/ncu_metrics.py REP substr [substr ...] (exact-ish filters)."""
import csv, subprocess, sys
rep, pats = sys.argv[1], sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    for k, unit, v in zip(h, u, r):
        if any(p in k for p in pats):
            print(f"{k:100s} {v} {unit}")
