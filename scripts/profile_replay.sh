#!/bin/bash
TAG=${1:-r01}
mkdir -p gpurun_out
python scripts/replay_bench.py 1e7 > gpurun_out/${TAG}_replay_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_replay_plain.log; exit 1; }
cat gpurun_out/${TAG}_replay_plain.log | grep -v Warn
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_replay_launches.csv python scripts/replay_bench.py 1e7 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hist_kernel|shower_kernel|eval_fn_kernel|sample_kernel|minmax_kernel" -s 12 -c 24 -f -o gpurun_out/${TAG}_replay_prof python scripts/replay_bench.py 1e7 > /dev/null 2>&1
echo done
