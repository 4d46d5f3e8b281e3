"""A script written against the REFERENCE's import paths, run on the B200 path without touching its import lines:
``install_as`` registers the mirror under ``jetmontecarlo`` (and the catalog under ``file_management``).
Body: critical + subsequent emission at fixed coupling, the flow of
examples/tests/multipleEmissions_tests/LL/test_critsub_sudakov.py:151-258 without the plotting, then the result is
stored in and read back from the reference's file catalog."""
import tempfile

import jetmontecarlo_b200
jetmontecarlo_b200.install_as("jetmontecarlo")

# ---- from here on: the reference's own lines ----
import numpy as np
from jetmontecarlo.montecarlo.integrator import *
from jetmontecarlo.numerics.radiators.samplers import *
from jetmontecarlo.numerics.observables import *
from jetmontecarlo.numerics.radiators.generation import *
from jetmontecarlo.analytics.radiators.fixedcoupling import *
from file_management.data_management import save_new_data, load_data
from file_management import catalog_utils

catalog_utils.set_output_root(tempfile.mkdtemp())          # (the reference writes below ./output)
NUM_SAMPLES, NUM_BINS, z_c, beta, eps, jet_type = int(1e7), 100, .1, 2., 1e-10, 'quark'

crit_sampler = criticalSampler('log', zc=z_c, epsilon=eps)
crit_sampler.generateSamples(NUM_SAMPLES)
samples = crit_sampler.getSamples()
z_crit, theta_crit = samples[:, 0], samples[:, 1]
sub_sampler = ungroomedSampler('log', epsilon=eps)
sub_sampler.generateSamples(NUM_SAMPLES)
c_sub = sub_sampler.getSamples()[:, 0]

obs = C_groomed(z_crit, theta_crit, z_c, beta, z_pre=0., f=1., acc='LL') + c_sub
test_int = integrator()
test_int.setLastBinBndCondition([1., 'minus'])
test_int.setBins(NUM_BINS, obs, 'log')
weights = criticalEmissionWeight(z_crit, theta_crit, z_c, jet_type, fixedcoupling=True) \
    * subPDFAnalytic_fc_LL(c_sub, beta, jet_type=jet_type)
jacs = np.array(crit_sampler.jacobians) * np.array(sub_sampler.jacobians)
test_int.setDensity(obs, weights * jacs, np.array(crit_sampler.area) * np.array(sub_sampler.area))
test_int.integrate()

params = {'z_cut': z_c, 'beta': beta, 'jet type': jet_type, 'number of samples': NUM_SAMPLES}
save_new_data({'bins': test_int.bins, 'integral': test_int.integral, 'integral error': test_int.integralErr},
              'numerical integral', 'critical sudakov function', params, '.npz')
stored = load_data('numerical integral', 'critical sudakov function', params)
print("cdf at C = 1e-6, 1e-3, 1e-1:", np.interp([1e-6, 1e-3, 1e-1], stored['bins'], stored['integral']))
print("integrator class:", type(test_int).__module__)
