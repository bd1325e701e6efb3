#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q -s > gpurun_out/r_large_tests.log 2>&1
echo "large tests rc=$?" >> gpurun_out/r_large_tests.log
grep "^\[large\|passed\|failed\|rc=\|Error\|assert\|^E " gpurun_out/r_large_tests.log | tail -30 | cut -c1-300
timeout -s KILL 300 python tools/bench_large.py > gpurun_out/r_bench_large.json 2>gpurun_out/r_bench_large.err; tail -c 900 gpurun_out/r_bench_large.json
