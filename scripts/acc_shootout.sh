#!/bin/bash
# Shared-memory accumulation back ends and launch shapes, measured on the real kernels (run under gpurun).
# Each arg is "ACC_MODE:HIST_MODE:HIST_UNROLL[:FAST_SHIFT[:MIN_BLOCKS[:PAIRS[:HIST_SHIFTS]]]]" (HIST_SHIFTS: comma list
# of maximal log2(copies) tried at run time); rebuilds the library and times every fused kind + jmc_hist.
for v in "$@"; do
  IFS=: read acc hm hu fs mb pr hs <<< "$v"
  export JMC_ACC_MODE=$acc JMC_HIST_MODE=$hm JMC_HIST_UNROLL=$hu
  [ -n "$fs" ] && export JMC_FAST_SPREAD_SHIFT=$fs || unset JMC_FAST_SPREAD_SHIFT
  [ -n "$mb" ] && export JMC_FAST_MIN_BLOCKS=$mb || unset JMC_FAST_MIN_BLOCKS
  [ -n "$pr" ] && export JMC_FAST_PAIRS=$pr || unset JMC_FAST_PAIRS
  python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== ACC_MODE=$acc HIST_MODE=$hm HIST_UNROLL=$hu FAST_SHIFT=${fs:-5} MIN_BLOCKS=${mb:-3} PAIRS=${pr:-1}"
  python scripts/kind_bench.py 2>&1 | tail -10
  for h in $(echo ${hs:-5} | tr ',' ' '); do
    echo "--- JMC_HIST_SPREAD_SHIFT=$h"
    JMC_HIST_SPREAD_SHIFT=$h python scripts/hist_bench.py 5e8 2>&1 | tail -8 | head -6
  done
done
unset JMC_ACC_MODE JMC_HIST_MODE JMC_HIST_UNROLL JMC_FAST_SPREAD_SHIFT JMC_FAST_MIN_BLOCKS JMC_FAST_PAIRS
python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1
