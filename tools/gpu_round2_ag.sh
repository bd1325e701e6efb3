#!/bin/bash
# round 2, session 3: large-map pass -- rank update pipelined through shared memory at rank >= 16 (accumulator loads before the operand
# staging and 4 resident CTAs below), first-touch scan with independent loads, rows / columns of Sigma_0 staged before the update chain:
# parity tests, bench (known and unknown correspondence), launch list of the bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q > gpurun_out/ag_tests.log 2>&1
echo "large tests rc=$?"; tail -3 gpurun_out/ag_tests.log
timeout -s KILL 600 python tools/bench_large.py > gpurun_out/ag_bench_large.json 2> gpurun_out/ag_bench_large.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/ag_bench_large.json').read().strip().splitlines()[-1])
print([(r['m'], round(r['ms_per_scan'],4), round(r['scans_per_s']), round(r['frac_of_hbm'],3)) for r in d['per_m']])
print([(r['m'], round(r['ms_per_scan'],4), r['ids_matching_truth']) for r in d.get('unknown_association_4000_candidates')])
P
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_large -c 600 --csv --log-file gpurun_out/launches_large_r02.csv python tools/bench_large.py > gpurun_out/ah_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_large_r02.csv')) if len(r)>5 and r[0].isdigit()]
seq=[(r[4].split('(')[0].replace('nuslam::',''), float(r[-1])) for r in rows]
agg=collections.defaultdict(list)
for n,v in seq: agg[n].append(v)
for n,v in agg.items(): print(n, len(v), 'min', min(v), 'med', sorted(v)[len(v)//2], 'max', max(v))
P
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_large_rank_update_pipe -s 4 -c 1 -f -o gpurun_out/prof_large_rank_pipe python tools/bench_large.py > gpurun_out/ag_ncu.log 2>&1
tail -1 gpurun_out/ag_ncu.log | cut -c1-120
