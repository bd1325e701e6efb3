#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 50 --warmup 5 > gpurun_out/s_bench_n${NG:-2}.json 2> gpurun_out/s_bench_n${NG:-2}.err
echo "n2 rc=$?"; tail -3 gpurun_out/s_bench_n${NG:-2}.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open("gpurun_out/s_bench_n"+__import__("os").environ.get("NG","2")+".json").read().strip().splitlines()[-1])
print('N=${NG:-2}', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('copy_ceiling'), 'assoc', d['extra']['assoc']['value'], d['extra']['assoc']['ms_per_step'])
P
timeout -s KILL 300 python -m pytest tests/test_shard.py -m gpu -x -q 2>&1 | tail -3
