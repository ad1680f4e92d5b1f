#!/bin/bash
# Standard GPU pass (run under gpurun): parity tests, smoke, bench, ncu launch list.
# Usage: tools/gpu_check.sh [tests|bench|ncu|all]...   outputs under gpurun_out/
mkdir -p gpurun_out
what="${*:-all}"
has() { [[ " $what " == *" $1 "* || " $what " == *" all "* ]]; }
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
if has tests; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/bench.log
  tail -c 600 gpurun_out/bench.log
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python bench.py --profile-one > gpurun_out/ncu_launches.log 2>&1
  echo "ncu exit $?" | tee -a gpurun_out/ncu_launches.log
fi
tail -5 gpurun_out/pytest_gpu.log 2>/dev/null
