#!/bin/bash
# All ncu evidence of a round in one gpurun call.  Reports are summarised ON THE BOX (scripts/ncu_summary.py) and only
# the text comes back (gpurun merges at most 64 MiB); KEEP names the one .ncu-rep to keep.
TAG=${1:-r02z}; KEEP=${2:-$TAG}
run() {  # tag, samples per launch, profile command...
  local tag=$1 per=$2; shift 2
  "$@" > gpurun_out/${tag}_run.log 2>&1
  if [ -f gpurun_out/${tag}_prof.ncu-rep ]; then
    python scripts/ncu_summary.py gpurun_out/${tag}_prof.ncu-rep $per hot > gpurun_out/${tag}_summary.txt 2>&1
    [ "$tag" = "$KEEP" ] || rm -f gpurun_out/${tag}_prof.ncu-rep
  fi
}
run ${TAG} 1e9 scripts/profile.sh ${TAG}
run ${TAG}_counted 1e9 scripts/profile.sh ${TAG}_counted --counts all
run ${TAG}_fc 1e9 scripts/profile.sh ${TAG}_fc --variant c3_fc
run ${TAG}_shower 2e7 scripts/profile_shower.sh ${TAG}_shower
python scripts/hist_bench.py 5e8 > gpurun_out/${TAG}_hist_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_hist_prof \
    python scripts/hist_bench.py 5e8 > gpurun_out/${TAG}_hist_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/${TAG}_hist_prof.ncu-rep 5e8 hot > gpurun_out/${TAG}_hist_summary.txt 2>&1
rm -f gpurun_out/${TAG}_hist_prof.ncu-rep
ls -la gpurun_out | tail -30
