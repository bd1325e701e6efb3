#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res2a_step -s 6 -c 1 -f -o gpurun_out/prof_res2a_v1 python tools/bench_assoc.py 131072 > gpurun_out/v_ncu.log 2>&1
tail -1 gpurun_out/v_ncu.log | cut -c1-150
