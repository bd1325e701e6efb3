#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/bench_large.py > gpurun_out/r_bench_large.json 2>gpurun_out/r_bench_large.err
python -c "
import json; d=json.load(open('gpurun_out/r_bench_large.json')); print([(p['m'],round(p['ms_per_scan'],3),round(p['frac_of_hbm'],3)) for p in d['per_m']]); print(d['unknown_association_4000_candidates'])"
timeout -s KILL 900 python -m pytest tests/test_large_gpu.py -m gpu -x -q > gpurun_out/r_large_tests.log 2>&1; tail -2 gpurun_out/r_large_tests.log
