#!/bin/bash
# round 2, session 3: association kernel with the rank-2 fragment update by plain FMAs instead of half-empty DMMAs
mkdir -p gpurun_out
NUSLAM_B200_LIB=build/variants/lib_adfma.so timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/ae_tests.log 2>&1
echo "adfma: ekf+world tests rc=$?"; tail -3 gpurun_out/ae_tests.log
for lib in shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_adfma.so shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_adfma.so; do
NUSLAM_B200_LIB=$lib python tools/bench_assoc.py 131072 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  assoc $lib', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
NUSLAM_B200_LIB=build/variants/lib_adfma.so timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-200
