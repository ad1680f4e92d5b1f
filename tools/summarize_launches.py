#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
    name = re.sub(r"\(.*", "", r[kn])
    name = re.sub(r"^void ", "", name)[:90]
    d[name][0] += 1
    d[name][1] += v
tot = sum(v[1] for v in d.values())
print(f"# {path}: {sum(v[0] for v in d.values())} launches, {tot:.1f} us total (per-launch times are cold-cache, serialised)")
print(f"{'us':>10} {'share':>6} {'n':>4}  kernel")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:4d}  {k}")
