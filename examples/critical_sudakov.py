"""Critical-emission Sudakov factor at fixed coupling: Monte Carlo cdf against the closed form.
ref: examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py:124-220"""
import numpy as np

from jetmontecarlo_b200.analytics.sudakov_factors.fixedcoupling import critSudakov_fc_LL
from jetmontecarlo_b200.montecarlo.integrator import integrator
from jetmontecarlo_b200.numerics.observables import C_groomed
from jetmontecarlo_b200.numerics.radiators.samplers import criticalSampler
from jetmontecarlo_b200.numerics.weights import criticalEmissionWeight

Z_CUT, BETA, EPSILON, NUM_SAMPLES, NUM_BINS = .1, 2., 1e-10, int(1e7), 100

for jet_type in ('quark', 'gluon'):
    crit_sampler = criticalSampler('log', zc=Z_CUT, epsilon=EPSILON)
    crit_sampler.generateSamples(NUM_SAMPLES)
    samples = crit_sampler.getSamples()
    z, theta = samples[:, 0], samples[:, 1]
    obs = C_groomed(z, theta, Z_CUT, BETA)
    weights = criticalEmissionWeight(z, theta, Z_CUT, jet_type, fixedcoupling=True)

    it = integrator()
    it.setLastBinBndCondition([1., 'minus'])
    it.setBins(NUM_BINS, obs, 'log')
    it.setDensity(obs, weights * np.array(crit_sampler.jacobians), crit_sampler.area)
    it.integrate()

    exact = np.array(critSudakov_fc_LL(it.bins, Z_CUT, BETA, jet_type), dtype=float)
    worst = np.abs(np.asarray(it.integral) - exact).max()
    print("%-5s  max |MC cdf - closed form| = %.2e   (largest MC error %.1e)" % (jet_type, worst, it.integralErr.max()))
