#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "pair_kernel or fast_step_shapes or free_running or teacher" > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c_tests.log
tail -8 gpurun_out/c_tests.log
( timeout -s KILL 600 bash tools/bench_variants.sh build/variants/lib_p8.so ) > gpurun_out/c_variants.log 2>&1
cat gpurun_out/c_variants.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_ekf_pair_step -s 4 -c 1 -f -o gpurun_out/prof_pair_v2 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/c_ncu.log 2>&1
tail -2 gpurun_out/c_ncu.log
