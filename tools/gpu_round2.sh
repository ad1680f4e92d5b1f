#!/bin/bash
# Round-2 GPU pass (run under gpurun): stages selectable by name.
#   tests smoke bench refarm refgpu sanitize ncu_fps ncu_pointconv ncu_linear ncu_knn launches train
mkdir -p gpurun_out
what="${*:-tests smoke bench refgpu}"
has() { [[ " $what " == *" $1 "* ]]; }
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
if has tests; then
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if has smoke; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/bench.log
  python tools/bench_line.py gpurun_out/bench.log 2>/dev/null | head -40
fi
if has refarm; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "refarm exit $?" | tee -a gpurun_out/bench_ref.log
fi
if has refgpu; then
  timeout 900 python tools/bench_reference_gpu.py > gpurun_out/reference_gpu.log 2>&1; echo "refgpu exit $?" | tee -a gpurun_out/reference_gpu.log
  grep -v '^{' gpurun_out/reference_gpu.log | tail -40
fi
if has sanitize; then
  tools/sanitize.sh memcheck racecheck
fi
NCU="ncu --set full --clock-control none --import-source on"
if has ncu_fps; then
  timeout 600 $NCU -k regex:fps_cluster -s 1 -c 1 -f -o gpurun_out/r02_fps python tools/prof_ops.py fps > gpurun_out/ncu_fps.log 2>&1; echo "ncu_fps exit $?"
fi
if has ncu_pointconv; then
  timeout 600 $NCU -k regex:PointConvProducer -s 1 -c 1 -f -o gpurun_out/r02_pointconv python tools/prof_ops.py pointconv > gpurun_out/ncu_pc.log 2>&1; echo "ncu_pointconv exit $?"
fi
if has ncu_linear; then
  timeout 600 $NCU -k regex:tc_gemm_kernel -s 3 -c 3 -f -o gpurun_out/r02_linear python tools/prof_ops.py linear > gpurun_out/ncu_lin.log 2>&1; echo "ncu_linear exit $?"
fi
if has ncu_knn; then
  timeout 600 $NCU -k regex:knn_bf_kernel -s 3 -c 3 -f -o gpurun_out/r02_knn python tools/prof_ops.py knn > gpurun_out/ncu_knn.log 2>&1; echo "ncu_knn exit $?"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python bench.py --profile-one > gpurun_out/ncu_launches.log 2>&1
  echo "launches exit $?"
  python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -30 gpurun_out/launches_summary.txt
fi
if has train; then
  timeout 900 python bench.py --workload kd_train --steps 10 --warmup 3 > gpurun_out/train1.log 2>&1; echo "train exit $?" | tee -a gpurun_out/train1.log
  tail -c 700 gpurun_out/train1.log
fi
