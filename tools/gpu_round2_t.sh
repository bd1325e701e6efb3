#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_scan_gpu.py tests/test_world_gpu.py tests/test_facade.py -m gpu -x -q -s > gpurun_out/t_scan_tests.log 2>&1
echo "scan tests rc=$?" >> gpurun_out/t_scan_tests.log
grep "scan_detect noise\|passed\|failed\|rc=\|^E " gpurun_out/t_scan_tests.log | tail -12 | cut -c1-220
timeout -s KILL 200 python tools/bench_scan.py > gpurun_out/t_bench_scan.log 2>&1
tail -1 gpurun_out/t_bench_scan.log | cut -c1-420
