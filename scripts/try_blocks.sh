#!/bin/bash
# rebuild the library on the GPU box with different register caps and time the headline kernel
for mb in "$@"; do
  JMC_FAST_MIN_BLOCKS=$mb python jetmontecarlo_b200/csrc/build.py --force > /dev/null 2>&1
  for v in c3_rc c3_fc; do
    JMC_FAST_MIN_BLOCKS=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minblocks $mb', d['config']['variant'], '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], d['config']['cdf_first_bin'])"
  done
done
