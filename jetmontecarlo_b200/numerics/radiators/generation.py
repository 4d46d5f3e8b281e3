"""Numerical radiators over parameter grids -- host mirror of
jetmontecarlo/numerics/radiators/generation.py (gen_numerical_radiator :36-140, gen_pre_num_rad :143-250,
gen_crit_sub_num_rad :253-400).  These are the reference's heaviest callers of the hot path (BASELINE config 5:
num_bins grid points x the full sample set each).  Here the sample set is uploaded once and stays in HBM;
every grid point is a handful of kernel launches on device-resident arrays, nothing per sample returns to the host.
"""
import numpy as np

from ...montecarlo.integrator import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..observables import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..weights import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ...analytics.radiators.running_coupling import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ...analytics.radiators.fixedcoupling import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ...utils.interpolation import is_monotonic_arr, is_monotonic_func, monotonic_domain  # noqa: F401,F403  (re-exported as the reference's module does)
from ... import _device as D
from ...montecarlo.integrator import integrator, integrator_2d
from ...utils.interpolation import get_2d_interpolation, lin_log_mixed_list
from ..observables import C_ungroomed, C_ungroomed_max
from ..weights import radiatorWeight
from ..._arrayfn import eval_fn
from ...analytics.qcd_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ... import _lib

local_verbose = 0


def _device_columns(sampler_obj):
    """(z, theta, jac) of a 2-D sampler as CUDA tensors, uploading host samples once if needed"""
    dev = getattr(sampler_obj, 'device_samples', None)
    samples = sampler_obj.getSamples()
    if dev is None or dev.shape[0] != len(samples):
        dev = D.to_device(np.asarray(samples, dtype=np.float64))
    jac = getattr(sampler_obj, 'device_jacobians', None)
    if jac is None or jac.shape[0] != len(samples):
        jac = D.to_device(np.asarray(sampler_obj.jacobians, dtype=np.float64)[-len(samples):])
    return dev[:, 0], dev[:, 1], jac          # column views: the kernels read them with their stride


def gen_numerical_radiator(rad_sampler, emission_type, jet_type='quark', obs_accuracy='LL', splitfn_accuracy='LL',
                           beta=None, bin_space='lin', fixed_coupling=True, num_bins=100):
    """Radiator depending on one parameter.  ref: numerics/radiators/generation.py:36-140"""
    assert (beta is None and not emission_type == 'sub')\
        or (beta is not None and emission_type == 'sub'),\
        str(emission_type) + ' emissions must have a valid beta value.'
    rad_integrator = integrator()
    rad_integrator.setLastBinBndCondition([0., 'plus'])
    rad_integrator.setMonotone()
    z, theta, jacs = _device_columns(rad_sampler)
    if emission_type == 'crit':
        obs = theta
    elif emission_type == 'sub':
        obs = C_ungroomed(z, theta, beta, acc=obs_accuracy)
    else:
        raise AssertionError("Invalid emission type. Emission must be "
                             + "'crit', 'sub', or 'pre', but is"
                             + str(emission_type))
    # like the reference, the bin range comes from the whole (N,2) sample array, not from obs (:103)
    rad_integrator.setBins(num_bins, rad_sampler.getSamples(), bin_space)
    weights = radiatorWeight(z, theta, jet_type, fixedcoupling=fixed_coupling, acc=splitfn_accuracy)
    rad_integrator.setDensity(obs, eval_fn(_lib.FN_MUL, [weights, jacs]), rad_sampler.area)
    rad_integrator.integrate()
    rad_integrator.makeInterpolatingFn(verbose=local_verbose)
    return rad_integrator.interpFn, rad_integrator.montecarlo_data_dict()


def gen_pre_num_rad(rad_sampler, crit_rad_sampler, jet_type='quark', splitfn_accuracy='LL', bin_space='lin',
                    fixed_coupling=True, num_bins=100):
    """Pre-critical radiator R(z_pre, theta_crit).  ref: numerics/radiators/generation.py:143-250"""
    rad_integrator = integrator_2d()
    rad_integrator.setLastBinBndCondition([0., 'plus'])
    theta_crit = _device_columns(crit_rad_sampler)[1]
    z_em = D.to_device(np.asarray(rad_sampler.getSamples(), dtype=np.float64))
    obs = [z_em, theta_crit]
    rad_integrator.setBins(num_bins, obs, bin_space)
    weights = radiatorWeight(z_em, theta_crit, jet_type, fixedcoupling=fixed_coupling, acc=splitfn_accuracy)
    if bin_space == 'lin':
        area = rad_sampler.zc
    else:
        area = np.log(1. / rad_sampler.epsilon) * np.log(1. / crit_rad_sampler.epsilon)
        weights = eval_fn(_lib.FN_MUL, [weights, eval_fn(_lib.FN_MUL, [z_em, theta_crit])])
    rad_integrator.setDensity(obs, weights, area)
    rad_integrator.integrate()
    rad_integrator.makeInterpolatingFn()
    unbounded_interp_function = rad_integrator.interpFn

    def bounded_interp_function(x, theta):
        return unbounded_interp_function((x, theta)) * (x >= 0) * (x <= rad_sampler.zc) * (theta >= 0)

    rad_integrator.interpFn = bounded_interp_function
    return rad_integrator.interpFn, rad_integrator.montecarlo_data_dict()


def gen_crit_sub_num_rad(rad_sampler, jet_type='quark', obs_accuracy='LL', splitfn_accuracy='LL', beta=2.,
                         epsilon=1e-15, bin_space='log', fixed_coupling=True, num_bins=1000):
    """Critical-subsequent radiator on a grid of theta_crit.  ref: numerics/radiators/generation.py:253-400.
    The loop body per grid point (rescale the angles, min/max, histogram, cumulative sum) is the reference's,
    on device-resident samples."""
    rads_all, rad_error_all, xs_all, thetas_all = [], [], [], []
    theta_calc_list = lin_log_mixed_list(epsilon, 1., num_bins)
    z_em, theta_samp, jacs_dev = _device_columns(rad_sampler)
    for theta_crit in theta_calc_list:
        rad_integrator = integrator()
        rad_integrator.setLastBinBndCondition([0., 'plus'])
        rad_integrator.setMonotone()
        jacs, area = jacs_dev, rad_sampler.area
        theta_em = eval_fn(_lib.FN_MUL, [theta_samp, float(theta_crit)])
        if bin_space == 'lin':
            area = area * theta_crit
        if bin_space == 'log':
            jacs = eval_fn(_lib.FN_MUL, [jacs_dev, float(theta_crit)])
        obs = C_ungroomed(z_em, theta_em, beta, acc=obs_accuracy)
        rad_integrator.setBins(num_bins, obs, bin_space, min_log_bin=theta_crit ** beta * 1e-15)
        weights = radiatorWeight(z_em, theta_em, jet_type, fixedcoupling=fixed_coupling, acc=splitfn_accuracy)
        rad_integrator.setDensity(obs, eval_fn(_lib.FN_MUL, [weights, jacs]), area)
        rad_integrator.integrate()
        xs = rad_integrator.bins
        rads_all.append(np.array(rad_integrator.integral))
        rad_error_all.append(np.array(rad_integrator.integralErr))
        xs_all.append(np.array(xs))
        thetas_all.append(np.ones(len(xs)) * theta_crit)
    xs_all, thetas_all, rads_all = np.array(xs_all), np.array(thetas_all), np.array(rads_all)
    unbounded_interp_function = get_2d_interpolation(xs_all.flatten(), thetas_all.flatten(), rads_all.flatten(),
                                                     interpolation_method='nearest')

    def bounded_interp_function(x, theta):
        return unbounded_interp_function(x, theta) * (x >= 0) * (theta >= 0) *\
            (x <= C_ungroomed_max(beta, radius=theta, acc=obs_accuracy))

    data_dict = {'xs': xs_all.flatten(), 'ys': thetas_all.flatten(), 'integral': rads_all.flatten()}
    return bounded_interp_function, data_dict


def gen_crit_sub_num_rad_fused(num_samples, jet_type='quark', obs_accuracy='LL', splitfn_accuracy='LL', beta=2.,
                               epsilon=1e-15, bin_space='log', fixed_coupling=True, num_bins=1000, radius=1.,
                               sample_epsilon=None, seed=0, precision='f32', group=None):
    """``gen_crit_sub_num_rad`` without a stored sample set: the whole theta_crit grid runs as two batched launches
    (min/max for the reference's data-dependent bins, then the histograms), each grid point regenerating the same
    ``num_samples`` ungroomed-sampler draws from the counter RNG.  Same return value as the reference function
    (numerics/radiators/generation.py:253-400)."""
    from ... import fused
    theta_calc_list = lin_log_mixed_list(epsilon, 1., num_bins)
    eps = epsilon if sample_epsilon is None else sample_epsilon
    igs = [fused.Integrand.radiator(bin_space, eps, radius=radius, beta=beta, jet_type=jet_type,
                                    fixedcoupling=fixed_coupling, obs_acc=obs_accuracy, split_acc=splitfn_accuracy,
                                    theta_scale=float(t)) for t in theta_calc_list]
    res = fused.integrate_batch(igs, num_samples, numbins=num_bins, binspacing=bin_space,
                                min_log_bins=theta_calc_list ** beta * 1e-15, seed=seed, precision=precision,
                                last_bc=[0., 'plus'], group=group)
    xs_all = res['bins']
    thetas_all = np.repeat(theta_calc_list[:, None], xs_all.shape[1], axis=1)
    rads_all = res['integral']
    unbounded_interp_function = get_2d_interpolation(xs_all.flatten(), thetas_all.flatten(), rads_all.flatten(),
                                                     interpolation_method='nearest')

    def bounded_interp_function(x, theta):
        return unbounded_interp_function(x, theta) * (x >= 0) * (theta >= 0) *\
            (x <= C_ungroomed_max(beta, radius=theta, acc=obs_accuracy))

    data_dict = {'xs': xs_all.flatten(), 'ys': thetas_all.flatten(), 'integral': rads_all.flatten(),
                 'integral error': res['integralErr']}
    return bounded_interp_function, data_dict
