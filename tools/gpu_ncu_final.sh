#!/bin/bash
# ncu --set full captures of the two kernels whose shared-memory layouts changed last (run under gpurun, one GPU).
# stages: costvol pointconv
mkdir -p gpurun_out
what="${*:-costvol pointconv}"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
if [[ " $what " == *" costvol "* ]]; then
  timeout 400 $NCU -k regex:CostVolPair -c 1 -f -o gpurun_out/r02j_costvol_pair python tools/prof_ops.py costvol > gpurun_out/ncu_cvp.log 2>&1; echo "ncu costvol pair exit $?"
fi
if [[ " $what " == *" pointconv "* ]]; then
  timeout 400 $NCU -k 'regex:tc_gemm_kernel.*PointConvProducer' -c 1 -f -o gpurun_out/r02j_pointconv python tools/prof_ops.py pointconv > gpurun_out/ncu_pc.log 2>&1; echo "ncu pointconv exit $?"
fi
ls -la gpurun_out/r02j_*.ncu-rep
