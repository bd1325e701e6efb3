#!/bin/bash
# resident pair kernel, variants after the calls left the kernel: pipelined state chain, chunks of 2, wraps as branches
mkdir -p gpurun_out
tools/bench_variants.sh build/variants/lib_pipe.so build/variants/lib_ch2.so build/variants/lib_branchy.so shermbot-navigation_b200/libnuslam_b200.so 2>&1 | tee gpurun_out/y_variants.log
