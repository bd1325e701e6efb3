#!/bin/bash
mkdir -p gpurun_out
NUSLAM_KERNEL=res2 NUSLAM_B200_LIB=$1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res2_step -s 4 -c 1 -f -o gpurun_out/$2 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/o_ncu.log 2>&1
tail -2 gpurun_out/o_ncu.log | cut -c1-200
