#!/bin/bash
# round 2, default = resident pair kernel: full GPU suite, bench line, launch list, ncu of the headline kernel
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p_tests_all.log 2>&1
echo "all tests rc=$?" >> gpurun_out/p_tests_all.log
tail -4 gpurun_out/p_tests_all.log
timeout -s KILL 900 python bench.py --steps 100 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/p_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/p_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_copy_ceiling'), d['details']['adversarial_ring'], d['roofline']['kernel'][:40])
P
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err; tail -c 400 gpurun_out/p_bench_ref.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/p_ncu_launches.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_res2_step -s 4 -c 1 -f -o gpurun_out/prof_res2_v3 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/p_ncu.log 2>&1
tail -1 gpurun_out/p_ncu.log | cut -c1-200
