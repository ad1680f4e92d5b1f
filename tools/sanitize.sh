#!/bin/bash
# compute-sanitizer over one small launch of each hand-synchronised kernel (run under gpurun).
# Usage: tools/sanitize.sh [memcheck racecheck synccheck]   -> gpurun_out/sanitize_<tool>.log
mkdir -p gpurun_out
tools="${*:-memcheck racecheck}"
filter='regex=tc_gemm_kernel|fps_cluster_kernel|fps_smem_kernel|knn_bf_kernel|spatial_sort_kernel|splitk_reduce|costvol_prep|pointconv_weightnet|interp3|linear_simt|knn_kernel'
for t in $tools; do
  timeout 1500 compute-sanitizer --tool "$t" --kernel-name "$filter" --launch-timeout 0 --error-exitcode 3 \
      --print-limit 40 python tools/sanitize_small.py > "gpurun_out/sanitize_$t.log" 2>&1
  echo "compute-sanitizer $t exit $?" | tee -a "gpurun_out/sanitize_$t.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done" "gpurun_out/sanitize_$t.log" | tail -3
done
