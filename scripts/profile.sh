#!/bin/bash
# ncu evidence for the fused kernel (run under gpurun, one GPU).  Usage: scripts/profile.sh <tag> [bench args]
# 1. plain run (must exit 0), 2. launch list with device times, 3. one --set full capture of the top kernel.
set -u
TAG=${1:-r02}; shift || true
ARGS="--steps 2 --warmup 3 --jobs 4 --no-cpu-baseline --no-sub-results $*"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:run_fast_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
