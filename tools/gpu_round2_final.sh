#!/bin/bash
# round 2, final tree: smoke(), full GPU suite, bench line with every extra, reference arm, launch list, ncu --set full of the headline,
# scan and association kernels
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/f_smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests_all.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/f_tests_all.log
timeout -s KILL 900 python bench.py --steps 100 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/f_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/f_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_copy_ceiling'), d['gpu_launches'], d['roofline']['launches_per_step'], d['details'].get('adversarial_ring'))
for k,v in d['extra'].items(): print(k, json.dumps(v)[:300])
P
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; tail -c 300 gpurun_out/f_bench_ref.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/f_ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res2_step -s 4 -c 1 -f -o gpurun_out/prof_res2_final \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/f_ncu.log 2>&1
tail -1 gpurun_out/f_ncu.log | cut -c1-160
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_scan_moment -s 2 -c 1 -f -o gpurun_out/prof_scan_moment_final python tools/bench_scan.py > gpurun_out/f_ncu_scan.log 2>&1
tail -1 gpurun_out/f_ncu_scan.log | cut -c1-160
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_ekf_fast_step -s 6 -c 1 -f -o gpurun_out/prof_assoc_final python tools/bench_assoc.py 131072 > gpurun_out/f_ncu_assoc.log 2>&1
tail -1 gpurun_out/f_ncu_assoc.log | cut -c1-160
timeout -s KILL 300 python tools/bench_closed_loop.py > gpurun_out/f_closed_loop.json 2> gpurun_out/f_closed_loop.err
tail -c 400 gpurun_out/f_closed_loop.json
