#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_world_gpu.py -m gpu -x -q > gpurun_out/ak_tests.log 2>&1
echo "world tests rc=$?"; tail -2 gpurun_out/ak_tests.log
timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-420
