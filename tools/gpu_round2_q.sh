#!/bin/bash
# round 2 evidence: ncu --set full of the scan, large-map and association kernels (one launch each), closed-loop bench
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_moment -s 2 -c 1 -f -o gpurun_out/prof_scan_moment_v2 python tools/bench_scan.py > gpurun_out/q_ncu_scan.log 2>&1
tail -1 gpurun_out/q_ncu_scan.log | cut -c1-150
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_rank_update -s 8 -c 1 -f -o gpurun_out/prof_large_rank_v1 python tools/bench_large.py > gpurun_out/q_ncu_large.log 2>&1
tail -1 gpurun_out/q_ncu_large.log | cut -c1-150
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_updates_coop -s 8 -c 1 -f -o gpurun_out/prof_large_coop_v1 python tools/bench_large.py > gpurun_out/q_ncu_large2.log 2>&1
tail -1 gpurun_out/q_ncu_large2.log | cut -c1-150
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_ekf_fast_step -s 6 -c 1 -f -o gpurun_out/prof_assoc_v1 python tools/bench_assoc.py 131072 > gpurun_out/q_ncu_assoc.log 2>&1
tail -1 gpurun_out/q_ncu_assoc.log | cut -c1-150
timeout -s KILL 300 python tools/bench_closed_loop.py > gpurun_out/q_closed_loop.json 2> gpurun_out/q_closed_loop.err
tail -c 600 gpurun_out/q_closed_loop.json
