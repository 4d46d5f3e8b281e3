#!/bin/bash
# Experiments on the fused FP32 kernel: each arg is a list of NAME=VALUE build switches separated by commas
# ("JMC_ACC_MODE=9,JMC_FAST_PAIRS=2"); rebuilds jmc_fast.cu only and prints scripts/kind_bench.py (run under gpurun).
export JMC_BUILD_ONLY=jmc_fast.cu
for v in "$@"; do
  for kv in $(echo $v | tr ',' ' '); do export $kv; done
  python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "=== $v"
  python scripts/kind_bench.py 2>&1 | tail -10 | cut -c1-62
  for kv in $(echo $v | tr ',' ' '); do unset ${kv%%=*}; done
done
python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1
