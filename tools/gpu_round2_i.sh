#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 300 bash tools/bench_variants.sh build/variants/lib_st.so ) > gpurun_out/i_variants.log 2>&1
cat gpurun_out/i_variants.log
NUSLAM_B200_LIB=build/variants/lib_st.so timeout -s KILL 200 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "kernels_agree or fast_step_shapes" > gpurun_out/i_tests_st.log 2>&1
tail -2 gpurun_out/i_tests_st.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/i_tests_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/i_tests_all.log
tail -5 gpurun_out/i_tests_all.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_scan_moment -s 2 -c 1 -f -o gpurun_out/prof_scan_moment_v1 python tools/bench_scan.py > gpurun_out/i_ncu.log 2>&1
tail -2 gpurun_out/i_ncu.log
