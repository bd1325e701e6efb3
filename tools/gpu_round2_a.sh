#!/bin/bash
# first GPU check of the pair kernel: parity tests, then A/B timing of the variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout -s KILL 300 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "pair_kernel" > gpurun_out/a_pair_test.log 2>&1
echo "pair test rc=$?" >> gpurun_out/a_pair_test.log
tail -5 gpurun_out/a_pair_test.log
timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py -m gpu -q -s > gpurun_out/a_ekf_tests.log 2>&1
echo "ekf tests rc=$?" >> gpurun_out/a_ekf_tests.log
tail -15 gpurun_out/a_ekf_tests.log
( timeout -s KILL 600 bash tools/bench_variants.sh build/variants/lib_p11.so build/variants/lib_p10.so build/variants/lib_p9.so
  NUSLAM_PAIR=0 timeout -s KILL 300 bash tools/bench_variants.sh ) > gpurun_out/a_variants.log 2>&1
cat gpurun_out/a_variants.log
