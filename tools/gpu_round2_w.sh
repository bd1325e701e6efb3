#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_updates_coop -s 32 -c 1 -f -o gpurun_out/prof_large_coop_m12 python tools/bench_large.py > gpurun_out/w_ncu.log 2>&1
tail -1 gpurun_out/w_ncu.log | cut -c1-120
