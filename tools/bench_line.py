#!/usr/bin/env python
"""Print the headline fields of the last JSON line of a bench log."""
import json, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.log"
l = [x for x in open(path) if x.startswith("{")][-1]
d = json.loads(l)
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f}  roofline frac {d['roofline']['frac']:.4f}")
for k, v in d.get("kernels", {}).items():
    print(f"  {k:18s} {v.get('sec', 0) * 1e6:8.1f} us  {v.get('gbs', 0):8.1f} GB/s  {v.get('tflops', '')}")
