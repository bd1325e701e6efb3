#!/bin/bash
mkdir -p gpurun_out
NUSLAM_KERNEL=res timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res_step -s 4 -c 1 -f -o gpurun_out/prof_res_v1 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/l_ncu.log 2>&1
tail -2 gpurun_out/l_ncu.log | cut -c1-300
NUSLAM_KERNEL=res NUSLAM_B200_LIB=build/variants/lib_res4.so timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res_step -s 4 -c 1 -f -o gpurun_out/prof_res4_v1 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/l_ncu4.log 2>&1
tail -2 gpurun_out/l_ncu4.log | cut -c1-300
