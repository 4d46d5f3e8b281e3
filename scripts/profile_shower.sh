#!/bin/bash
# ncu capture of the veto-shower kernel (MLL quark, FP32, Soft Drop two-emission ECF), run under gpurun
set -u
TAG=${1:-r02_shower}
cat > gpurun_out/_shower_job.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import gpu_helpers as H
from jetmontecarlo_b200.fused import Integrand
edges = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
ig = Integrand.shower(beta=2., jet_type='quark', acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL', n_emissions=2)
for k in range(3):
    H.run(ig, 20_000_000, k, edges, 'f32')
print("done")
PY
python gpurun_out/_shower_job.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain failed"; tail -3 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:shower_fast_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof \
    python gpurun_out/_shower_job.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "shower capture rc=$?"
