"""Sampling primitives -- host mirror of jetmontecarlo/utils/montecarlo_utils.py.

``getLinSample`` / ``getLogSample`` / ``getLogSample_zerobin`` (:12-90) are the scalar building blocks of the
reference's Python sampling loop.  The samplers of this package never call them (they draw whole arrays in one
kernel); they are kept, drawing from the same Philox stream sequence, so that scripts using them keep running.
``inverse_transform_samples`` (:377-399) samples from a tabulated cdf on the GPU (jmc_inverse_transform).
``samples_from_cdf`` (:93-375) takes the cdf as a function: it is tabulated on the reference's probe points and
sampled on the GPU through the same kernel (the reference first tries a ``pynverse`` root finder and takes this
tabulated route whenever that refuses the cdf).
"""
import ctypes as C

import numpy as np

from .. import _device as D
from .. import _lib
from .._seed import next_key


def _uniforms(n, k=1):
    D.require_cuda()
    out = D.empty((n, k))
    _lib.check(_lib.lib.jmc_uniforms(n, next_key(), 0, k, D.ptr(out), D.stream()))
    return D.to_host(out)


def getLinSample(sample_min, sample_max):
    """ref: utils/montecarlo_utils.py:12-28"""
    return sample_min + float(_uniforms(1)[0, 0]) * (sample_max - sample_min)


def getLogSample(sample_min, sample_max, epsilon=1e-8):
    """ref: utils/montecarlo_utils.py:31-55"""
    logsample = (np.log(epsilon) * float(_uniforms(1)[0, 0]) + np.log(sample_max - sample_min))
    return sample_min + np.exp(logsample)


def getLogSample_zerobin(sample_min, sample_max, cum_dist, epsilon=1e-8):
    """ref: utils/montecarlo_utils.py:58-90"""
    u = _uniforms(1, 2)[0]
    if u[0] < cum_dist(epsilon):
        return 1e-50
    logsample = (np.log(epsilon) * float(u[1]) + np.log(sample_max - sample_min))
    return sample_min + np.exp(logsample)


def inverse_transform_samples(cdf_vals, x_vals, num_samples, seed=None):
    """Samples of the pdf whose cdf is tabulated as (x_vals, cdf_vals): linear interpolation of the inverse cdf
    at uniform random numbers, extrapolating beyond the table like interp1d(fill_value='extrapolate').
    ref: utils/montecarlo_utils.py:377-399"""
    D.require_cuda()
    cdf_vals, x_vals = np.asarray(cdf_vals, dtype=np.float64), np.asarray(x_vals, dtype=np.float64)
    order = np.argsort(cdf_vals, kind="mergesort")          # interp1d sorts its x (= cdf) axis
    c, x = D.to_device(cdf_vals[order]), D.to_device(x_vals[order])
    out = D.empty((int(num_samples),))
    _lib.check(_lib.lib.jmc_inverse_transform(D.ptr(c), D.ptr(x), int(c.numel()), int(num_samples),
                                              next_key() if seed is None else int(seed), 0, D.ptr(out), D.stream()))
    return D.to_host(out)


def _cdf_probe_points(domain):
    """ref: utils/montecarlo_utils.py:147-154"""
    from .interpolation import lin_log_mixed_list
    if domain[0] < 0:
        return np.linspace(domain[0], domain[1], 1000)
    if domain[0] == 0:
        return np.array([0, *lin_log_mixed_list(domain[0] + 1e-20, domain[1], 1000)])
    return lin_log_mixed_list(domain[0], domain[1], 1000)


def samples_from_cdf(cdf, num_samples, domain=None, catch_turnpoint=False, backup_cdf=None, force_monotone=False,
                     verbose=5, seed=None):
    """Samples (and unit weights) from a cdf given as a function on ``domain``, by the inverse-transform method.

    The reference first hands the cdf to pynverse's root finder and, when that refuses it, tabulates it on ~1000
    linearly + logarithmically spaced probe points and works through a list of known shapes.  Here the cdf is
    always tabulated (one vectorised call of ``cdf``), the same list is walked on the host (O(1000)), and the samples
    are drawn on the GPU (``inverse_transform_samples``); for a smooth monotone cdf this differs from exact
    inversion by the linear-interpolation error of the table.  In order:
      monotone table -> samples of the table;  cdf starting at 1 -> zero bin (all samples 0);
      cdf below 1e-20 everywhere -> overflow bin (all samples domain[1]);  below 1e-20 up to a point and monotone
      after it -> that part of the domain is dropped;  ``catch_turnpoint``: cdf reaching 1 and turning back -> the
      domain is cut at the turning point;  ``force_monotone``: table entries after which the cdf decreases are
      dropped;  otherwise ValueError (a ``backup_cdf`` needs the root finder and is not supported).
    ref: utils/montecarlo_utils.py:93-375"""
    assert domain is not None, "samples_from_cdf needs the domain of the cdf"
    pnts = _cdf_probe_points(domain)
    with np.errstate(all='ignore'):
        cdf_vals = np.nan_to_num(cdf(pnts))
    monotone = cdf_vals[:-1] <= cdf_vals[1:]
    retry = dict(catch_turnpoint=catch_turnpoint, backup_cdf=backup_cdf, force_monotone=force_monotone,
                 verbose=verbose, seed=seed)
    if monotone.all():
        samples = inverse_transform_samples(cdf_vals, pnts, num_samples, seed=seed)
        return samples, np.ones_like(samples)
    bad_low = pnts[:-1][~monotone]
    with np.errstate(all='ignore'):
        bad_cdf_low = cdf(np.unique(bad_low))
    if all(cdf_vals == 1.):
        if verbose > 4:
            print("CDF always 1. Returning zero bin.")
        return np.zeros(num_samples), np.ones(num_samples)
    if bad_low[0] == pnts[0] and bad_cdf_low[0] == 1:
        if verbose > 1:
            print("Found non-monotone cdf behavior at minimum value of domain. Drawing from zero bin.")
        return np.zeros(num_samples), np.ones(num_samples)
    cdf_threshold = 1e-20
    if all(cdf_vals < cdf_threshold):
        if verbose > 4:
            print("CDF always negligible. Returning highest point in domain.")
        samples = np.full(num_samples, domain[1])
        return samples, np.ones_like(samples)
    for i, (cdf_val, point) in enumerate(zip(cdf_vals, pnts)):
        if cdf_val > cdf_threshold:
            if not all(monotone[i:]):
                break
            if verbose > 4:
                print("CDF always negligible up to a certain point. Removing 'bad' part from the domain.")
            return samples_from_cdf(cdf, num_samples, domain=[point, domain[1]], **retry)
    if bad_cdf_low[0] == 1 and catch_turnpoint:
        if verbose > 1:
            print("Found turning point of CDF. Adjusting sampling.")
        return samples_from_cdf(cdf, num_samples, domain=[domain[0], bad_low[0]], **retry)
    if force_monotone:
        if verbose > 1:
            print("Found non-monotone cdf behavior. Forcing monotonic behavior.")
        samples = inverse_transform_samples(cdf_vals[:-1][monotone], pnts[:-1][monotone], num_samples, seed=seed)
        return samples, np.ones_like(samples)
    raise ValueError("samples_from_cdf: the cdf is not monotone on [%g, %g] (it decreases after x = %s ...) and "
                     "neither catch_turnpoint nor force_monotone resolves it"
                     % (domain[0], domain[1], bad_low[:3]))
