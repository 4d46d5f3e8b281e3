#!/bin/bash
# compute-sanitizer passes over the kernels with shared-memory histograms (run under gpurun, one GPU, in a call of
# its own: never together with ncu).  memcheck + racecheck on a small job of every fused kind, the replay histogram
# and the shower; the jobs are small because the tools slow kernels down by 10-100x.
set -u
mkdir -p gpurun_out
cat > gpurun_out/_sanitize_job.py <<'PY'
import numpy as np
from jetmontecarlo_b200 import _lib, fused
from jetmontecarlo_b200.fused import Integrand
from jetmontecarlo_b200.montecarlo.integrator import integrator
from jetmontecarlo_b200.utils import partonshower_utils as PS
edges = np.logspace(-11, np.log10(.5), 100)
jobs = [Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False, obs_acc='MLL'),
        Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=True, obs_acc='MLL'),
        Integrand.critical('log', .1, 1e-10, fixedcoupling=False, nan_to_num=True),
        Integrand.radiator('log', 1e-10, beta=2., jet_type='gluon', fixedcoupling=False, obs_acc='MLL', split_acc='MLL'),
        Integrand.pre_crit_sub('log', .1, 1e-10, nan_to_num=True)]
for ig in jobs:
    for prec in ('f32', 'f64'):
        bc = [0., 'plus'] if ig.kind == _lib.KIND_RADIATOR else [1., 'minus']
        fused.integrate(ig, 300_000, edges=edges, seed=3, precision=prec, last_bc=bc)
    fused.integrate(ig, 100_000, numbins=50, binspacing='log', seed=3, precision='f32', last_bc=[1., 'minus'])
rng = np.random.RandomState(0)
it = integrator()
obs, w = rng.rand(200_000), rng.rand(200_000)
it.setBins(64, obs, 'lin'); it.setLastBinBndCondition([1., 'minus']); it.setDensity(obs, w, 1.); it.integrate()
jets = PS.gen_jets(20_000, beta=2., jet_type='quark', acc='MLL', verbose=0, seed=1)
PS.getECFs_softdrop(jets, .1, beta=2., obs_acc='MLL', n_emissions=2)
print("jobs done")
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python gpurun_out/_sanitize_job.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
