#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 600 bash tools/bench_variants.sh build/variants/lib_p8s2.so build/variants/lib_p8s1.so build/variants/lib_p12s1.so ) > gpurun_out/f_variants.log 2>&1
cat gpurun_out/f_variants.log
NUSLAM_B200_LIB=build/variants/lib_t8.so NUSLAM_FAST_CTAS_PER_SM=8 NUSLAM_FILTERS_PER_WARP=2 timeout -s KILL 200 python tools/fast_timing.py > gpurun_out/f_timing.log 2>&1
cat gpurun_out/f_timing.log
timeout -s KILL 900 python -m pytest tests/test_scan_gpu.py tests/test_facade.py tests/test_world_gpu.py -m gpu -x -q -s > gpurun_out/f_scan_tests.log 2>&1
echo "scan tests rc=$?" >> gpurun_out/f_scan_tests.log
tail -25 gpurun_out/f_scan_tests.log
timeout -s KILL 300 python tools/bench_scan.py > gpurun_out/f_bench_scan.log 2>&1
tail -3 gpurun_out/f_bench_scan.log
