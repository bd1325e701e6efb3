#!/bin/bash
# round 2, session 3: association distance written out on the structure of the division-free rows (49 operations instead of 70 FMAs)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/ad_tests.log 2>&1
echo "ekf+world tests rc=$?"; tail -3 gpurun_out/ad_tests.log
for k in fast res2a; do
NUSLAM_KERNEL=$k python tools/bench_assoc.py 131072 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  assoc $k', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
NUSLAM_KERNEL=res2a timeout -s KILL 600 python -m pytest tests/test_ekf_gpu.py -m gpu -x -q -k "assoc or config4 or unknown" 2>&1 | tail -2
timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-200
