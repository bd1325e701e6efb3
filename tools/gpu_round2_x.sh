#!/bin/bash
# final check of the tree: smoke(), full GPU suite, short bench
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/x_smoke.log
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_tests_all.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/x_tests_all.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/x_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'], d['roofline']['launches_per_step'])
P
