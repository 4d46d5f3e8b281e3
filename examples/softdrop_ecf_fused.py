"""Soft Drop (z_cut = 0.1, beta_sd = 0) critical x subsequent MLL ECF distribution, 1e9 samples, fused path:
samples are generated, weighted and binned in one kernel and never stored (BASELINE config 3)."""
import numpy as np

from jetmontecarlo_b200 import fused
from jetmontecarlo_b200.analytics.radiators.soft_drop import softdrop_sudakov_fc

edges = np.logspace(-11, np.log10(.5), 100)
for fixed in (False, True):
    ig = fused.Integrand.crit_sub('log', z_cut=.1, epsilon=1e-10, beta=2., jet_type='quark', fixedcoupling=fixed,
                                  obs_acc='MLL')
    it = fused.integrate(ig, 10 ** 9, edges=edges, seed=1, precision='f32', last_bc=[1., 'minus'])
    cdf = np.asarray(it.integral)
    print("%s coupling: cdf(first edge) = %.4f +- %.4f, cdf monotone: %s"
          % ("fixed" if fixed else "running", cdf[0], it.integralErr[0], bool(np.all(np.diff(cdf) >= -1e-12))))
    if fixed:
        # the LL Soft Drop Sudakov factor is the curve this two-emission integrand approximates at small C
        print("   LL Soft Drop Sudakov at C = 1e-3, 1e-2, 1e-1:", softdrop_sudakov_fc(np.array([1e-3, 1e-2, 1e-1]), .1, 2.))
        print("   Monte Carlo cdf there:                       ", np.interp([1e-3, 1e-2, 1e-1], edges, cdf))
