#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ekf_gpu.py tests/test_world_gpu.py -m gpu -x -q -s -k "unknown_association or association_free_running or config4 or scan_step or closed_loop or replay" > gpurun_out/u_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/u_tests.log
grep "^\[\|passed\|failed\|rc=\|^E  " gpurun_out/u_tests.log | tail -24 | cut -c1-260
for k in res2 fast; do NUSLAM_KERNEL=$k timeout -s KILL 300 python tools/bench_assoc.py 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$k', round(d['ms_per_step'],3), 'ms per 1Mi-filter step', round(d['roofline']['frac'],4), d['stats']['bad_status'], d['stats'].get('ids_differing_from_truth_last_step'))"; done
