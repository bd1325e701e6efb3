#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q > gpurun_out/ai_tests.log 2>&1
echo "large tests rc=$?"; tail -2 gpurun_out/ai_tests.log
for r in 16 0 1000; do
NUSLAM_LARGE_PIPE_MIN_RANK=$r timeout -s KILL 600 python tools/bench_large.py > gpurun_out/ai_bench_large.json 2> gpurun_out/ai_bench_large.err; echo "min rank $r bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/ai_bench_large.json').read().strip().splitlines()[-1])
print([(r['m'], round(r['ms_per_scan'],4), round(r['scans_per_s']), round(r['frac_of_hbm'],3)) for r in d['per_m']])
P
done
