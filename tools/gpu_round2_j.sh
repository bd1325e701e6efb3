#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_scan_gpu.py tests/test_world_gpu.py -m gpu -x -q -s > gpurun_out/j_scan_tests.log 2>&1
echo "scan tests rc=$?" >> gpurun_out/j_scan_tests.log
grep "scan_detect noise\|passed\|failed\|rc=" gpurun_out/j_scan_tests.log | tail -12
timeout -s KILL 200 python tools/bench_scan.py > gpurun_out/j_bench_scan.log 2>&1
tail -2 gpurun_out/j_bench_scan.log | cut -c1-700
timeout -s KILL 900 python bench.py --steps 100 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/j_bench.json
tail -5 gpurun_out/j_bench.err
