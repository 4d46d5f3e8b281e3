#!/bin/bash
# rebuild with different compile-time choices and time the headline kernel: each arg is "MINBLOCKS:UNROLL:PIN[:PHILOX_SPLIT]"
for v in "$@"; do
  IFS=: read mb un pin pipe <<< "$v"; pipe=${pipe:-0}
  JMC_FAST_MIN_BLOCKS=$mb JMC_FAST_UNROLL=$un JMC_FAST_PIN_KEYS=$pin JMC_PHILOX_SPLIT=$pipe python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  for var in c3_rc c3_fc; do
    JMC_FAST_MIN_BLOCKS=$mb JMC_FAST_UNROLL=$un JMC_FAST_PIN_KEYS=$pin JMC_PHILOX_SPLIT=$pipe python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant $var 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks/unroll/pin $v', d['config']['variant'], '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], d['config']['cdf_first_bin'])"
  done
done
