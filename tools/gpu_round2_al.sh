#!/bin/bash
# round 2, session 3: scan kernel with the CTA-level fit phase (examined clusters of the CTA's four scans fitted together)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/al_tests.log 2>&1
echo "scan+world tests rc=$?"; tail -3 gpurun_out/al_tests.log
for k in 1 2; do
timeout -s KILL 300 python tools/bench_scan.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  scan', d.get('ms_per_pass'), d.get('value'), d.get('scans_rerun_in_oracle_order'))"
done
timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-330
