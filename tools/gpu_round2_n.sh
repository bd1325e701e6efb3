#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -s -k "kernels_agree" > gpurun_out/n_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/n_tests.log
grep "res2 vs fast\|passed\|failed\|rc=\|Error\|assert" gpurun_out/n_tests.log | tail -12
( NUSLAM_KERNEL=res2 timeout -s KILL 400 bash tools/bench_variants.sh "$@" 2>&1 | sed "s/^/res2: /" ) > gpurun_out/n_variants.log 2>&1
cat gpurun_out/n_variants.log
