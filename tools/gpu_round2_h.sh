#!/bin/bash
mkdir -p gpurun_out
for cfg in "one 1" "one 256" "edge 0" "one 4096"; do
  timeout -s KILL 60 python tools/dbg_scan.py $cfg 2>&1 | tail -8
  echo "rc=$? ($cfg)"
done > gpurun_out/h_dbg.log 2>&1
cat gpurun_out/h_dbg.log
timeout -s KILL 300 python -m pytest tests/test_scan_gpu.py tests/test_facade.py -m gpu -x -q -s > gpurun_out/h_scan_tests.log 2>&1
echo "scan tests rc=$?" >> gpurun_out/h_scan_tests.log
tail -25 gpurun_out/h_scan_tests.log
timeout -s KILL 200 python tools/bench_scan.py > gpurun_out/h_bench_scan.log 2>&1
tail -3 gpurun_out/h_bench_scan.log
