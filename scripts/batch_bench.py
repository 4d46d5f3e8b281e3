"""BASELINE config 5: theta_crit grid of the critical-subsequent radiator, one launch per pass (jmc_run_batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jetmontecarlo_b200 import fused
from jetmontecarlo_b200.fused import Integrand
from jetmontecarlo_b200.utils.interpolation import lin_log_mixed_list
n = 5_000_000
for fc, acc in ((True, 'LL'), (False, 'MLL')):
    for npts, nbins in ((500, 500), (2000, 500), (5000, 500), (500, 5000), (5000, 5000)):
        grid = lin_log_mixed_list(1e-10, 1., npts)
        igs = [Integrand.radiator('log', 1e-10, beta=2., jet_type='quark', fixedcoupling=fc, obs_acc=acc, split_acc=acc,
                                  theta_scale=float(t)) for t in grid]
        kw = dict(numbins=nbins, binspacing='log', min_log_bins=grid ** 2. * 1e-15, seed=1, precision='f32', last_bc=[0., 'plus'])
        fused.integrate_batch(igs[:4], 100000, **dict(kw, min_log_bins=kw['min_log_bins'][:4]))
        torch.cuda.synchronize(); t = time.perf_counter()
        res = fused.integrate_batch(igs, n, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{'fc LL ' if fc else 'rc MLL'} {len(grid):5d} points x {n:.0e} samples x {nbins:5d} edges: {dt:7.3f} s  "
              f"{2*len(grid)*n/dt:.3e} sample-evals/s (min/max pass + histogram pass)")
