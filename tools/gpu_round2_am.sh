#!/bin/bash
# launch list of the closed-loop bench (per-kernel durations of a step)
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_closed_loop_r02.csv python tools/bench_closed_loop.py > gpurun_out/am_ncu.log 2>&1
python - <<'P'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_closed_loop_r02.csv')) if len(r)>5 and r[0].isdigit()]
seq=[(r[4].split('(')[0].replace('nuslam::','').replace('void ',''), float(r[-1])) for r in rows]
agg=collections.defaultdict(list)
for n,v in seq[len(seq)//2:]: agg[n].append(v)
for n,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{n[:60]:<60} n={len(v):4d} med={sorted(v)[len(v)//2]/1e3:8.1f} us  max={max(v)/1e3:8.1f}")
P
