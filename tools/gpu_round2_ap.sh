#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_rank_update_pipe -s 4 -c 1 -f -o gpurun_out/prof_large_rank_pipe2 python tools/bench_large.py > gpurun_out/ap_ncu.log 2>&1
tail -1 gpurun_out/ap_ncu.log | cut -c1-120
