#!/bin/bash
# bench every kernel-experiment build under build/variants/ (and the in-tree library) on the headline workload
for lib in shermbot-navigation_b200/libnuslam_b200.so "$@"; do
  NUSLAM_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 3 --e2e-repeats 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['ms_per_step']*1000,1),'us  frac', round(d['roofline']['frac'],4), 'bad', d['bad_filters'])"
done
