#!/bin/bash
mkdir -p gpurun_out
for cfg in "one 1" "one 8" "one 256" "edge 0" "one 4096"; do
  timeout -s KILL 60 python tools/dbg_scan.py $cfg 2>&1 | tail -12
  echo "rc=$? ($cfg)"
done > gpurun_out/g_dbg.log 2>&1
cat gpurun_out/g_dbg.log
for k in static pair fast; do NUSLAM_KERNEL=$k timeout -s KILL 200 bash tools/bench_variants.sh 2>&1 | sed "s/^/$k: /"; done > gpurun_out/g_kernels.log 2>&1
cat gpurun_out/g_kernels.log
timeout -s KILL 400 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "kernels_agree or fast_step_shapes or free_running or teacher or error_stats" > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -6 gpurun_out/g_tests.log
