#!/bin/bash
mkdir -p gpurun_out
NUSLAM_KERNEL=res timeout -s KILL 400 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -s -k "kernels_agree" > gpurun_out/k_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/k_tests.log
grep "res vs fast\|passed\|failed\|rc=\|Error\|assert" gpurun_out/k_tests.log | tail -12
( NUSLAM_KERNEL=fast timeout -s KILL 200 bash tools/bench_variants.sh 2>&1 | head -1 | sed "s/^/fast: /"
  NUSLAM_KERNEL=res timeout -s KILL 400 bash tools/bench_variants.sh "$@" 2>&1 | sed "s/^/res: /" ) > gpurun_out/k_variants.log 2>&1
cat gpurun_out/k_variants.log
