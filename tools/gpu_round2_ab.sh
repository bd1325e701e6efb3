#!/bin/bash
# round 2, session 3: two-translation-unit build (FAST kernels relocatable, the rest whole-program): tests, headline, config 4, scan, closed loop;
# ncu --set full of the lane-major scan kernel
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py tests/test_abi.py tests/test_scan_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
for lib in shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_nordc.so; do
  echo "== $lib"
  NUSLAM_B200_LIB=$lib python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  headline', round(d['ms_per_step']*1000,1),'us  frac', round(d['roofline']['frac'],4), 'bad', d['bad_filters'], 'launches', d['roofline']['launches_per_step'])"
  NUSLAM_B200_LIB=$lib python tools/bench_assoc.py 131072 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  assoc', d['ms_per_step'], d['value'])"
  NUSLAM_B200_LIB=$lib timeout -s KILL 300 python tools/bench_scan.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  scan', d.get('ms_per_pass'), d.get('value'), d.get('scans_rerun_in_oracle_order'), d.get('mean_clusters_per_scan'), d.get('mean_circles_per_scan'))"
  NUSLAM_B200_LIB=$lib timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/ab_bench.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_moment -s 2 -c 1 -f -o gpurun_out/prof_scan_moment_v4 python tools/bench_scan.py > gpurun_out/ab_ncu_scan.log 2>&1
tail -1 gpurun_out/ab_ncu_scan.log | cut -c1-150
