#!/bin/bash
# 2-GPU pass (run under `gpurun --gpus 2`): NCCL gradient-parity tests, inference and KD-training bench lines under torchrun.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_nccl.py -x -q -m gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest2 exit $?" | tee -a gpurun_out/pytest_2gpu.log
tail -3 gpurun_out/pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 exit $?"
grep '^{' gpurun_out/bench_2gpu.log | tail -1 | cut -c1-260
timeout 600 $TR bench.py --gpus 2 --workload kd_train --steps 10 --warmup 3 > gpurun_out/train_2gpu.log 2>&1; echo "train2 exit $?"
grep '^{' gpurun_out/train_2gpu.log | tail -1 | cut -c1-700
