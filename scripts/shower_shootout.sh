#!/bin/bash
# Grouping thresholds / chunk / occupancy of the FP32 shower kernel, measured (run under gpurun).
# Each arg is "GROUP:BIN_GROUP:CHUNK:MIN_BLOCKS[:TRIALS]"; rebuilds jmc_shower.cu only and times scripts/shower_bench.py.
export JMC_BUILD_ONLY=jmc_shower.cu
for v in "$@"; do
  IFS=: read g b c m t <<< "$v"
  export JMC_SHOWER_GROUP=$g JMC_SHOWER_BIN_GROUP=$b JMC_SHOWER_CHUNK=$c JMC_SHOWER_MIN_BLOCKS=$m JMC_SHOWER_TRIALS=${t:-2}
  python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== GROUP=$g BIN_GROUP=$b CHUNK=$c MIN_BLOCKS=$m TRIALS=${t:-2}"
  python scripts/shower_bench.py 2>&1 | head -3
done
unset JMC_SHOWER_GROUP JMC_SHOWER_BIN_GROUP JMC_SHOWER_CHUNK JMC_SHOWER_MIN_BLOCKS JMC_SHOWER_TRIALS
python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1
