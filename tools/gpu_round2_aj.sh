#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_world_scan -s 6 -c 1 -f -o gpurun_out/prof_world_scan python tools/bench_closed_loop.py > gpurun_out/aj_ncu.log 2>&1
tail -1 gpurun_out/aj_ncu.log | cut -c1-120
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_updates_coop -s 40 -c 1 -f -o gpurun_out/prof_large_coop_m12b python tools/bench_large.py > gpurun_out/aj_ncu2.log 2>&1
tail -1 gpurun_out/aj_ncu2.log | cut -c1-120
