#!/bin/bash
# round 2, session 3: device-side tail launch of the list kernel + call-free sincos / division in the FAST kernels
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py tests/test_abi.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/z_tests.log 2>&1
echo "ekf+abi+world tests rc=$?"; tail -3 gpurun_out/z_tests.log
for lib in shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_nordc.so shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_nordc.so; do
  NUSLAM_B200_LIB=$lib python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['ms_per_step']*1000,1),'us  frac', round(d['roofline']['frac'],4), 'bad', d['bad_filters'], 'launches', d['roofline']['launches_per_step'])"
  NUSLAM_B200_LIB=$lib python tools/bench_assoc.py 131072 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  assoc', d['ms_per_step'], d['value'])"
done 2>&1 | tee gpurun_out/z_bench.log
