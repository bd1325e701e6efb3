#!/bin/bash
# round 2, session 3: scan kernel after the Newton / pass-2 trims: parity tests, timing, closed loop
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_scan_gpu.py tests/test_world_gpu.py -m gpu -x -q > gpurun_out/ac_tests.log 2>&1
echo "scan+world tests rc=$?"; tail -3 gpurun_out/ac_tests.log
timeout -s KILL 300 python tools/bench_scan.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  scan', d.get('ms_per_pass'), d.get('value'), d.get('scans_rerun_in_oracle_order'), d['cpu_baseline']['value'])"
timeout -s KILL 300 python tools/bench_closed_loop.py 2>/dev/null | tail -1 | cut -c1-200
