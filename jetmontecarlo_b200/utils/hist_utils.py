"""``vals_to_pdf`` -- host mirror of jetmontecarlo/utils/hist_utils.py:94-144: weighted pdf of a sample on fixed
lin bins on [0,1] or log bins down to ``log_cutoff`` with a zero bin [1e-100, cutoff), normalised to unit
integral and, for log bins, multiplied by x ln 10 (density in log10 x).  The per-sample work (finite mask,
histogram) runs on the GPU (JMC_FN_MASK_FINITE, jmc_hist); the O(bins) normalisation is NumPy as in the reference.
"""
from warnings import warn

import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from .._arrayfn import eval_fn


def vals_to_pdf(vals, num_bins, weights=None, bin_space='log', log_cutoff=None, make_finite=True):
    """ref: utils/hist_utils.py:94-144"""
    v = D.to_device(vals).reshape(-1)
    w = D.to_device(np.ones(v.numel()) if weights is None else weights).reshape(-1)
    if make_finite:
        # a removed sample == a sample of weight 0 (the pdf is normalised by the histogram's own sum)
        w = eval_fn(_lib.FN_MASK_FINITE, [w, v])
        n_finite = C.c_uint64(0)
        _lib.check(_lib.lib.jmc_count_finite(D.ptr(v), v.numel(), C.byref(n_finite), D.stream()))
        if n_finite.value == 0:
            warn("No finite values in `vals`, the values given to "
                 "vals_to_pdf. Returning empty arrays.")
            return np.array([]), np.array([])
    if bin_space == 'lin':
        bins = np.linspace(0, 1, num_bins)
        xs = (bins[:-1] + bins[1:]) / 2
    elif bin_space == 'log':
        assert log_cutoff is not None,\
            "log cutoff required for logarithmic"\
            " bins (you could try, say, -10)."
        bins = np.logspace(np.log10(log_cutoff), 0, num_bins)
        bins = np.insert(bins, 0, 1e-100)  # zero bin
        xs = np.sqrt(bins[1:-1] * bins[2:])
        xs = np.insert(xs, 0, 1e-50)  # zero bin
    e = D.to_device(bins)
    nb = e.numel() - 1
    sw, sw2 = D.zeros((nb,)), D.zeros((nb,))
    cnt = D.zeros((nb,), dtype=torch.int64)
    _lib.check(_lib.lib.jmc_hist(D.ptr(v), D.ptr(w), v.numel(), D.ptr(e), nb + 1, D.ptr(sw), D.ptr(sw2), D.ptr(cnt),
                                 D.stream()))
    pdf = D.to_host(sw)
    dx = np.array(np.diff(bins), float)
    pdf = pdf / dx / pdf.sum()
    if bin_space == 'log':
        pdf = xs * pdf * np.log(10)
    return xs, pdf


# ---- O(bins) post-processing of finished histograms (host, as in the reference) --------------------------------
def histDerivative(hist, bins, giveHist=False, binInput='lin'):
    """Second-order finite-difference derivative of a histogram with uniform (or, ``binInput='log'``, log-uniform)
    bins: one-sided three-point stencils in the first and last bin, central differences between; for log bins the
    derivative is with respect to the variable itself, not its logarithm.  Returns an extrapolating interpolant
    over the bin centres (and the values there if ``giveHist``).  ref: utils/hist_utils.py:7-91"""
    from scipy import interpolate
    hist = np.asarray(hist)
    if binInput == 'log':
        bins = np.log(bins)
    width = bins[-1] - bins[-2]
    first = np.array([(-hist[2] + 4. * hist[1] - 3. * hist[0]) / (2. * width)])
    bulk = (hist[2:] - hist[:-2]) / (2. * width)
    last = np.array([(3. * hist[-1] - 4. * hist[-2] + hist[-3]) / (2. * width)])
    if binInput == 'lin':
        xs = (bins[1:] + bins[:-1]) / 2.
        deriv = np.concatenate((first, bulk, last))
    elif binInput == 'log':
        xs = np.exp((bins[1:] + bins[:-1]) / 2.)
        deriv = np.concatenate((first, bulk, last)) / xs
    else:
        raise AssertionError("The style of the input bins must be either 'lin' or 'log'.")
    interp = interpolate.interp1d(x=xs, y=deriv, fill_value="extrapolate")
    if giveHist:
        return interp, deriv
    return interp


def nan_info(array, name):
    """ref: utils/hist_utils.py:146-149"""
    print(f'nan {name}')
    print(int(np.isnan(np.asarray(array, dtype=float)).sum()))


def notfinite_info(array, name):
    """ref: utils/hist_utils.py:152-155"""
    print(f'notfinite {name}')
    print(int((~np.isfinite(np.asarray(array, dtype=float))).sum()))


def smooth_data(x, y):
    """Low-pass filter: real FFT coefficients whose square is below 5 % of the largest are dropped.
    ref: utils/hist_utils.py:158-173"""
    import scipy.fftpack
    assert len(x) == len(y), "x and y must be the same length."
    coeffs = scipy.fftpack.rfft(y)
    power = coeffs ** 2
    kept = np.where(power < power.max() * 5e-2, 0., coeffs)
    return scipy.fftpack.irfft(kept)
