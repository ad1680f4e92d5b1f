mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_layers_gpu.py -x -q -m gpu -k "column_blocks or concat_free or wide_linear" > gpurun_out/pytest_new.log 2>&1; echo "new exit $?" | tee -a gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
KDPC_CONCAT_FREE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_cat.log 2>&1; echo "bench0 exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_free.log 2>&1; echo "bench1 exit $?"
tail -c 300 gpurun_out/bench_cat.log; echo; tail -c 300 gpurun_out/bench_free.log
