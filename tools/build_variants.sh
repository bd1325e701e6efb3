#!/bin/bash
# build kernel-experiment variants of the library into build/variants/: tools/build_variants.sh name "-DFLAG=.. -DFLAG2=.." [name flags]...
# (the two-unit build of shermbot-navigation_b200/build.py); prints registers / spills of the kernels matching $KPAT
cd "$(dirname "$0")/.."
mkdir -p build/variants
KPAT=${KPAT:-k_ekf_res2_stepILi12}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  python shermbot-navigation_b200/build.py --variant $name $flags 2>&1 | grep -A2 "$KPAT" | grep "registers\|spill\|error" | sed "s/^/$name: /" &
done
wait
