#!/bin/bash
# large-map pipelined rank update: operands of 1 / 2 / 3 k-steps requested together
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q > gpurun_out/ai_tests.log 2>&1
echo "large tests rc=$?"; tail -2 gpurun_out/ai_tests.log
for lib in shermbot-navigation_b200/libnuslam_b200.so build/variants/lib_ks1.so build/variants/lib_ks3.so; do
NUSLAM_B200_LIB=$lib timeout -s KILL 600 python tools/bench_large.py > gpurun_out/ai_bench_large.json 2> gpurun_out/ai_bench_large.err; echo "$lib bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/ai_bench_large.json').read().strip().splitlines()[-1])
print([(r['m'], round(r['ms_per_scan'],4), round(r['scans_per_s']), round(r['frac_of_hbm'],3)) for r in d['per_m']])
P
done
