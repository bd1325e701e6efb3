#!/bin/bash
# driver-style short bench (K = 20, W = 5), twice, and a K = 100 line
mkdir -p gpurun_out
for k in 20 20 100; do
python bench.py --gpus 1 --steps $k --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 20 --e2e-repeats 1 2>gpurun_out/aq.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('K=$k', round(d['ms_per_step']*1000,2),'us  frac', round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], 'bad', d['bad_filters'], 'launches', d['gpu_launches'])"
done
tail -2 gpurun_out/aq.err
