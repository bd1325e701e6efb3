#!/bin/bash
# build kernel-experiment variants of the library into build/variants/: tools/build_variants.sh name "-DFLAG=.. -DFLAG2=.." [name flags]...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared --expt-relaxed-constexpr \
    -Iinclude -Ishermbot-navigation_b200/csrc $flags -o build/variants/lib_$name.so shermbot-navigation_b200/csrc/nuslam_b200.cu -Xptxas -v 2>&1 \
    | grep -A2 "k_ekf_pair_step\|k_ekf_fast_stepILi12ELb1ELb0" | grep "registers\|spill" | sed "s/^/$name: /" &
done
wait
