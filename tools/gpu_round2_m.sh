#!/bin/bash
mkdir -p gpurun_out
( NUSLAM_KERNEL=res timeout -s KILL 500 bash tools/bench_variants.sh "$@" 2>&1 | sed "s/^/res: /" ) > gpurun_out/m_variants.log 2>&1
cat gpurun_out/m_variants.log
