#!/bin/bash
mkdir -p gpurun_out
for v in 8 12; do
  echo "== pair kernel, $v CTAs/SM"
  NUSLAM_B200_LIB=build/variants/lib_t$v.so NUSLAM_FAST_CTAS_PER_SM=$v NUSLAM_FILTERS_PER_WARP=2 timeout -s KILL 200 python tools/fast_timing.py
done > gpurun_out/d_timing.log 2>&1
cat gpurun_out/d_timing.log
