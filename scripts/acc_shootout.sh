#!/bin/bash
# Shared-memory accumulation back ends, measured on the real kernels (run under gpurun):
# each arg is "ACC_MODE:HIST_MODE:HIST_UNROLL"; rebuilds the library and times every fused kind + jmc_hist.
for v in "$@"; do
  IFS=: read acc hm hu <<< "$v"
  export JMC_ACC_MODE=$acc JMC_HIST_MODE=$hm JMC_HIST_UNROLL=$hu
  python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== ACC_MODE=$acc HIST_MODE=$hm HIST_UNROLL=$hu"
  python scripts/kind_bench.py 2>&1 | tail -10
  python scripts/hist_bench.py 5e8 2>&1 | tail -10
done
unset JMC_ACC_MODE JMC_HIST_MODE JMC_HIST_UNROLL
python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1
