#!/bin/bash
# Pipeline parameter sweep (run under gpurun): SM cap of the forward graphs' persistent kernels, one / two forward streams.
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/sweep_$name.log 2>&1
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
lines = [l for l in open(f"gpurun_out/sweep_{name}.log") if l.startswith("{")]
if not lines:
    print(name, "FAILED"); sys.exit(0)
d = json.loads(lines[-1])
print(f"{name:14s} value {d['value']:8.1f}  e2e {d['e2e']['value']:8.1f}  ms {d['ms_per_step']:.3f}  sm_limit {d['details']['pipeline_sm_limit']}  streams {d['details']['pipeline_forward_streams']}")
PY
}
for lim in 116 124 132 140 148; do run sm$lim KDPC_SM_LIMIT=$lim; done
run single_fwd KDPC_DUAL_FORWARD=0
