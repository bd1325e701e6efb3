#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "pair_kernel or fast_step_shapes or free_running or teacher" > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e_tests.log
tail -4 gpurun_out/e_tests.log
( timeout -s KILL 600 bash tools/bench_variants.sh build/variants/lib_p8s2.so build/variants/lib_p8s1.so build/variants/lib_p12s1.so ) > gpurun_out/e_variants.log 2>&1
cat gpurun_out/e_variants.log
NUSLAM_B200_LIB=build/variants/lib_t8.so NUSLAM_FAST_CTAS_PER_SM=8 NUSLAM_FILTERS_PER_WARP=2 timeout -s KILL 200 python tools/fast_timing.py > gpurun_out/e_timing.log 2>&1
cat gpurun_out/e_timing.log
