#!/bin/bash
# round 2, session 3: lane-major clustering walk of the scan moment kernel -- parity tests, timing (8 / 6 resident CTAs), closed loop and
# world / strict timing of the relocatable-device-code build against the whole-program build
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/aa_tests.log 2>&1
echo "scan+world tests rc=$?"; tail -3 gpurun_out/aa_tests.log
for lib in shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_scan6.so build/variants/lib_nordc.so; do
  echo "== $lib"
  NUSLAM_B200_LIB=$lib timeout -s KILL 300 python tools/bench_scan.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  scan', d.get('ms_per_pass'), d.get('value'), d.get('scans_rerun_in_oracle_order'))"
  NUSLAM_B200_LIB=$lib timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-300
done 2>&1 | tee gpurun_out/aa_bench.log
