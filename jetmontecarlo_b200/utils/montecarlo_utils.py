"""Sampling primitives -- host mirror of jetmontecarlo/utils/montecarlo_utils.py.

``getLinSample`` / ``getLogSample`` / ``getLogSample_zerobin`` (:12-90) are the scalar building blocks of the
reference's Python sampling loop.  The samplers of this package never call them (they draw whole arrays in one
kernel); they are kept, drawing from the same Philox stream sequence, so that scripts using them keep running.
``inverse_transform_samples`` (:377-399) samples from a tabulated cdf on the GPU (jmc_inverse_transform).
``samples_from_cdf`` (:93-375) needs ``pynverse`` root finding and is not on the accelerated path.
"""
import ctypes as C

import numpy as np

from .. import _device as D
from .. import _lib
from ..montecarlo.sampler import next_key


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
