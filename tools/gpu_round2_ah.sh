#!/bin/bash
# launch list of the large-map bench (per-kernel durations of one pass at m = 1, 12, 32)
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_large -c 600 --csv --log-file gpurun_out/launches_large_r02.csv python tools/bench_large.py > gpurun_out/ah_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_large_r02.csv')) if len(r)>5 and r[0].isdigit()]
import collections
seq=[(r[4].split('(')[0].replace('nuslam::',''), float(r[-1])) for r in rows]
print(len(seq))
# print the last 40 launches (m = 32 section) and a window in the middle
for name,v in seq[-12:]: print(name, v)
agg=collections.defaultdict(list)
for n,v in seq: agg[n].append(v)
for n,v in agg.items(): print(n, len(v), 'min', min(v), 'med', sorted(v)[len(v)//2], 'max', max(v))
P
