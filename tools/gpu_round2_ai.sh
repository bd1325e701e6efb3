#!/bin/bash
# large-map pipelined rank update: operands staged in shared memory (3 CTAs per SM) against operands from L2 / L1 (4 CTAs)
mkdir -p gpurun_out
NUSLAM_B200_LIB=build/variants/lib_so1.so timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in c4 so1 so1pf0 c4 so1pf0; do
NUSLAM_B200_LIB=build/variants/lib_$v.so timeout -s KILL 600 python tools/bench_large.py > gpurun_out/ai_bench_large.json 2> gpurun_out/ai_bench_large.err; echo "$v bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/ai_bench_large.json').read().strip().splitlines()[-1])
print([(r['m'], round(r['ms_per_scan'],4), round(r['scans_per_s']), round(r['frac_of_hbm'],3)) for r in d['per_m']])
P
done
