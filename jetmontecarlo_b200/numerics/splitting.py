"""Normalised splitting functions -- host mirror of jetmontecarlo/numerics/splitting.py:16-60.

For every theta of a lin/log grid the reference runs a fresh ``integrate_1d`` of alpha(z, theta) P(z) over
z in (z_cut, 1/2) (a new Python-loop sampler per grid point).  Here each grid point is: draw on the GPU, evaluate
alpha * P in one kernel (JMC_FN_ALPHA_SPLITTING), histogram into the single bin, all on device-resident arrays.
"""
import numpy as np
from scipy import interpolate

from .. import _lib
from .._arrayfn import eval_fn
from ..analytics.qcd_utils import _jet, alpha_fixed, alpha_s, splittingFn
from ..montecarlo.integrator import integrator
from ..montecarlo.sampler import simpleSampler
from ..utils.interpolation import lin_log_mixed_list


def _integrate_alpha_p(theta, z_cut, jet_type, accuracy, fixed_coupling, bin_space, epsilon, num_samples):
    """integrate_1d(weight, [z_cut, 1/2], num_bins=2) of the reference (integrator.py:390-457), device resident"""
    this_sampler = simpleSampler(bin_space, epsilon=None if bin_space == 'lin' else epsilon, bounds=[z_cut, 1. / 2.])
    samples, jacs = this_sampler._draw(int(num_samples), this_sampler.__class__._next_key())
    this_integrator = integrator()
    this_integrator.setFirstBinBndCondition(0)
    this_integrator.setBins(2, samples, bin_space)
    weights = eval_fn(_lib.FN_ALPHA_SPLITTING, [samples, float(theta)], jet=_jet(jet_type),
                      acc=_lib.ACCS[accuracy], fixed_coupling=fixed_coupling)
    this_integrator.setDensity(samples, eval_fn(_lib.FN_MUL, [weights, jacs]), this_sampler.area)
    this_integrator.integrate()
    return this_integrator.integral[1]


def gen_normalized_splitting(num_samples, z_cut, jet_type='quark', accuracy='LL', fixed_coupling=True,
                             bin_space='lin', epsilon=1e-15, num_bins=100, verbose=1):
    """ref: numerics/splitting.py:16-60"""
    assert accuracy in ['LL', 'LL_nonsing', 'MLL'], "Invalid accuracy."
    theta_calc_list, norms = lin_log_mixed_list(epsilon, 1., num_bins), []
    for theta in theta_calc_list:
        norms.append(_integrate_alpha_p(theta, z_cut, jet_type, accuracy, fixed_coupling, bin_space, epsilon,
                                        num_samples))
    normalization = interpolate.interp1d(x=theta_calc_list, y=norms, fill_value="extrapolate")

    def normed_splitting_fn(z, theta):
        if fixed_coupling:
            alpha = alpha_fixed
        else:
            alpha = alpha_s(z, theta)
        splitfn = alpha * splittingFn(z, jet_type, accuracy) / normalization(theta)
        return splitfn * (z_cut < z) * (z < 1. / 2.)

    return normed_splitting_fn
