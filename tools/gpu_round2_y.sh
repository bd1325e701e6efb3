#!/bin/bash
# round 2, session 3: what-if builds of the resident pair kernel (timing only) + the pipelined state chain
mkdir -p gpurun_out
NUSLAM_B200_LIB=build/variants/lib_pipe.so timeout -s KILL 600 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "kernels_agree or golden or free_running" > gpurun_out/y_tests_pipe.log 2>&1
echo "pipe tests rc=$?"; tail -3 gpurun_out/y_tests_pipe.log
tools/bench_variants.sh build/variants/lib_base.so build/variants/lib_pipe.so build/variants/lib_ch2.so build/variants/lib_pipe2.so build/variants/lib_exp1.so build/variants/lib_exp2.so build/variants/lib_exp3.so build/variants/lib_exp5.so build/variants/lib_exp8.so 2>&1 | tee gpurun_out/y_variants.log
