#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_ekf_pair_step -s 4 -c 1 -f -o gpurun_out/prof_pair_v1 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
