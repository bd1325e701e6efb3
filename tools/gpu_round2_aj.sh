#!/bin/bash
# ncu --set full of the simulator's lidar kernel (final tree: two-phase scan)
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_world_scan -s 6 -c 1 -f -o gpurun_out/prof_world_scan2 python tools/bench_closed_loop.py > gpurun_out/aj_ncu.log 2>&1
tail -1 gpurun_out/aj_ncu.log | cut -c1-120
