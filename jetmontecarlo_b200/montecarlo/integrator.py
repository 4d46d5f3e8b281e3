"""Monte Carlo integrator -- host mirror of jetmontecarlo/montecarlo/integrator.py.

``integrator`` keeps the reference's methods, attributes, flags, assertion
messages and on-disk formats.  The numerics run in libjmc_b200.so:

  setBins      min / max of the observable       -> jmc_minmax (edges themselves
               are np.linspace / np.logspace on the host, as in the reference)
  setDensity   3 x np.histogram + normalisation  -> jmc_hist + jmc_finalize
  integrate    (reverse) cumulative sums         -> jmc_cumulative (block scan)

Observables and weights may be NumPy arrays (copied to the device) or CUDA
tensors (used in place).  Additive API for the fused path, where samples are
never materialised: ``setBinsFused`` / ``setDensityFused`` take an integrand
description (jetmontecarlo_b200.fused.Integrand) instead of arrays.
"""
import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from .sampler import simpleSampler
from ..utils.hist_utils import histDerivative, nan_info  # noqa: F401  (held by the reference's module as well)
from ..utils.interpolation import get_1d_interpolation, get_2d_interpolation  # noqa: F401

MIN_LOG_BIN = 1e-15


def _bc_args(first_bc, last_bc):
    """(bc_kind, bc_value) of the C ABI from the reference's two attributes."""
    use_last = (last_bc is not None) and first_bc is None
    use_first = (first_bc is not None) and last_bc is None
    ambiguous = (not (use_last or use_first)) or (use_last and use_first)
    assert not ambiguous, "Integration requires unambiguous boundary conditions"
    if use_first:
        return _lib.BC_FIRST, float(first_bc)
    assert last_bc[1] in ['plus', 'minus'], \
        "Integration requires lastBinBndCond (plus or minus)"
    return (_lib.BC_LAST_PLUS if last_bc[1] == 'plus' else _lib.BC_LAST_MINUS), float(last_bc[0])


class integrator():
    """Weighted histogram -> density +- error -> cumulative integral +- error.
    ref: montecarlo/integrator.py:21-386"""

    # ------------------
    # Integration Tools:
    # ------------------
    def _device_obs(self, observables):
        """The observable array on the device (CUDA tensors are used in place; a host array is uploaded every time it
        is passed: the caller may have changed it in place between setBins and setDensity, and nothing is pinned)."""
        return D.to_device(observables).reshape(-1)

    def _min_max(self, observables):
        dev = self._device_obs(observables)
        if dev.numel() == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        out = D.empty((2,))
        _lib.check(_lib.lib.jmc_minmax(D.ptr(dev), dev.numel(), D.ptr(out), D.stream()))
        lo, hi = D.to_host(out)
        return lo, hi

    def _edges_from_range(self, numbins, obs_min, obs_max, binspacing, min_log_bin):
        """ref: montecarlo/integrator.py:41-55"""
        if binspacing == 'lin':
            self.bins = np.linspace(min(0, obs_min), obs_max, numbins)
        elif binspacing == 'log':
            min_bin = np.log10(max(obs_min, min_log_bin))
            max_bin = np.log10(obs_max)
            self.bins = np.logspace(min_bin, max_bin, numbins)
        self.binspacing = binspacing
        if self.binspacing == 'lin':
            self.bin_midpoints = (self.bins[1:] + self.bins[:-1]) / 2.
        elif self.binspacing == 'log':
            self.bin_midpoints = np.exp((np.log(self.bins[1:]) + np.log(self.bins[:-1])) / 2.)
        self.hasBins = True

    def setBins(self, numbins, observables, binspacing, min_log_bin=MIN_LOG_BIN):
        """Sets the bin edges from the range of the observables; ``numbins`` is
        the number of EDGES.  ref: montecarlo/integrator.py:37-57"""
        obs_min, obs_max = self._min_max(observables)
        self._edges_from_range(numbins, obs_min, obs_max, binspacing, min_log_bin)

    def setLastBinBndCondition(self, bndCond):
        self.lastBinBndCond = bndCond
        self.useLastBinBndCond = True

    def setFirstBinBndCondition(self, bndCond):
        self.firstBinBndCond = bndCond
        self.useFirstBinBndCond = True

    def setMonotone(self, monotone=True):
        self.monotone = monotone

    # ------------------
    # Actual Integration:
    # ------------------
    def _density_from_hist(self, area, n_total):
        """jmc_finalize on the device histograms -> density, densityErr."""
        nb = len(self.bins) - 1
        dens, err = D.empty((nb,)), D.empty((nb,))
        _lib.check(_lib.lib.jmc_finalize(
            D.ptr(self._sum_w), D.ptr(self._sum_w2), D.ptr(self._count), D.ptr(self._edges_dev), nb + 1,
            float(area), int(n_total), _lib.BC_FIRST, 0., D.ptr(dens), D.ptr(err), None, None, D.stream()))
        self.density = D.to_host(dens)
        self.densityErr = D.to_host(err)
        self.hasMCDensity = True

    def _alloc_hist(self):
        edges = np.ascontiguousarray(self.bins, dtype=np.float64)
        assert edges.ndim == 1 and len(edges) >= 2, "Need bins to produce density histogram."
        nb = len(edges) - 1
        self._edges_dev = D.to_device(edges)
        self._sum_w, self._sum_w2 = D.zeros((nb,)), D.zeros((nb,))
        self._count = D.zeros((nb,), dtype=torch.int64)
        return nb

    def setDensity(self, observables, weights, area):
        """Histogram of weights and squared weights in the observable, normalised
        to a density with its statistical error.  ref: montecarlo/integrator.py:83-127"""
        assert self.hasBins, "Need bins to produce density histogram."
        assert area is not None, "Need a valid area to produce density histogram."
        obs = self._device_obs(observables)
        w = D.to_device(weights).reshape(-1)
        if w.numel() != obs.numel():
            raise ValueError("weights should have the same shape as a.")
        nb = self._alloc_hist()
        _lib.check(_lib.lib.jmc_hist(D.ptr(obs), D.ptr(w), obs.numel(), D.ptr(self._edges_dev), nb + 1,
                                     D.ptr(self._sum_w), D.ptr(self._sum_w2), D.ptr(self._count), D.stream()))
        self._density_from_hist(area, obs.numel())

    def integrate(self):
        """Cumulative integral of the density from the boundary condition.
        ref: montecarlo/integrator.py:129-202"""
        assert self.hasMCDensity, \
            "Need density function to perform integration"
        bc_kind, bc_value = _bc_args(self.firstBinBndCond, self.lastBinBndCond)
        self.useLastBinBndCond = bc_kind != _lib.BC_FIRST
        self.useFirstBinBndCond = bc_kind == _lib.BC_FIRST
        edges = D.to_device(np.asarray(self.bins, dtype=np.float64))
        nb = edges.numel() - 1
        dens = D.to_device(np.asarray(self.density, dtype=np.float64))
        err = D.to_device(np.asarray(self.densityErr, dtype=np.float64))
        integral, ierr = D.empty((nb + 1,)), D.empty((nb,))
        _lib.check(_lib.lib.jmc_cumulative(D.ptr(dens), D.ptr(err), D.ptr(edges), nb + 1, bc_kind, bc_value,
                                           D.ptr(integral), D.ptr(ierr), D.stream()))
        # the reference keeps a Python list of length numbins here
        self.integral = list(D.to_host(integral))
        self.integralErr = D.to_host(ierr)
        self.hasMCIntegral = True
        if self.monotone:
            arr = np.asarray(self.integral)
            is_monotone = ((arr[1:] <= arr[:-1]).all() or (arr[1:] >= arr[:-1]).all())
            assert is_monotone, "User asked for a monotone integral,"\
                " but the integral is not monotone!"

    # ------------------
    # Fused path (additive API)
    # ------------------
    def setBinsFused(self, numbins, integrand, num_samples, binspacing, seed, min_log_bin=MIN_LOG_BIN,
                     precision='f64', group=None):
        """setBins for a described integrand: a min/max pass that regenerates
        the samples from the counter RNG instead of storing them."""
        from .. import fused
        obs_min, obs_max = fused.observable_range(integrand, num_samples, seed, precision, group)
        self._edges_from_range(numbins, obs_min, obs_max, binspacing, min_log_bin)

    def setDensityFused(self, integrand, num_samples, seed, precision='f32', group=None):
        """setDensity for a described integrand: generate -> weight -> bin in
        one kernel; with ``torch.distributed`` initialised the samples are
        sharded by index over the ranks of ``group`` and the histograms are
        summed with one all-reduce."""
        from .. import fused
        assert self.hasBins, "Need bins to produce density histogram."
        self._alloc_hist()
        fused.accumulate(integrand, num_samples, seed, precision, self._edges_dev,
                         self._sum_w, self._sum_w2, self._count, group, edges_host=self.bins)
        self._density_from_hist(integrand.area, int(num_samples))

    def integrateFused(self, integrand, num_samples, seed, precision='f32', group=None):
        """setDensityFused + integrate in one pass over the device: the histograms go through ONE jmc_finalize
        (density, its error, cumulative integral, its error) into one buffer and come back in ONE copy, instead of
        density down / density up / integral down."""
        from .. import fused
        assert self.hasBins, "Need bins to produce density histogram."
        bc_kind, bc_value = _bc_args(self.firstBinBndCond, self.lastBinBndCond)
        nb = self._alloc_hist()
        fused.accumulate(integrand, num_samples, seed, precision, self._edges_dev,
                         self._sum_w, self._sum_w2, self._count, group, edges_host=self.bins)
        out = D.empty((4 * nb + 1,))
        base, step = out.data_ptr(), 8
        _lib.check(_lib.lib.jmc_finalize(
            D.ptr(self._sum_w), D.ptr(self._sum_w2), D.ptr(self._count), D.ptr(self._edges_dev), nb + 1,
            float(integrand.area), int(num_samples), bc_kind, bc_value, C.c_void_p(base), C.c_void_p(base + step * nb),
            C.c_void_p(base + step * 2 * nb), C.c_void_p(base + step * (3 * nb + 1)), D.stream()))
        host = D.to_host(out)
        self.density, self.densityErr = host[:nb].copy(), host[nb:2 * nb].copy()
        self.integral, self.integralErr = list(host[2 * nb:3 * nb + 1]), host[3 * nb + 1:].copy()
        self.useLastBinBndCond = bc_kind != _lib.BC_FIRST
        self.useFirstBinBndCond = bc_kind == _lib.BC_FIRST
        self.hasMCDensity = self.hasMCIntegral = True
        if self.monotone:
            arr = np.asarray(self.integral)
            assert ((arr[1:] <= arr[:-1]).all() or (arr[1:] >= arr[:-1]).all()), \
                "User asked for a monotone integral, but the integral is not monotone!"

    # ------------------
    # Saved data         (keys of the reference: integrator.py:204-224)
    # ------------------
    def montecarlo_data_dict(self, info=None):
        data_dict = {'bins': self.bins,
                     'bin midpoints': self.bin_midpoints,
                     'density': self.density,
                     'density error': self.densityErr,
                     'integral': self.integral,
                     'integral error': self.integralErr}
        if info is not None:
            data_dict['info'] = info
        return data_dict

    def save_montecarlo_data(self, filename, info=None):
        assert self.hasMCDensity and self.hasMCIntegral, \
            "Need density function and integral to save MC data"
        np.savez(filename, **self.montecarlo_data_dict(info=info))

    # ------------------
    # Interpolation (host, O(bins))
    # ------------------
    def makeInterpolatingFn(self, **kwargs):
        from ..utils.interpolation import get_1d_interpolation
        assert self.hasMCIntegral, \
            "Need MC integral to produce interpolation"
        self.interpFn = get_1d_interpolation(self.bins, self.integral, monotonic=self.monotone, **kwargs)
        self.hasInterpIntegral = True

    def makeInterpolatingDensity(self, binspacing, monotone=False):
        from ..utils.interpolation import get_1d_interpolation
        assert self.hasMCDensity, \
            "Need MC density to produce interpolation"
        bins = self.bins
        self.interpDensity = get_1d_interpolation(self.bin_midpoints, self.density, monotonic=False,
                                                  bounds=(bins[0], bins[-1]), bound_values=(0., 0.))
        self.hasInterpDensity = True

    def saveInterpolatingFn(self, fileName, save_integral=True):
        """dill pickle of the interpolating function, as the reference writes it.  ref: integrator.py:241-247"""
        import dill
        assert self.hasInterpIntegral, \
            "No interpolating function to save"
        with open(fileName, 'wb') as file:
            dill.dump(self.interpFn, file)

    def loadInterpolatingFn(self, fileName):
        """ref: integrator.py:249-254"""
        import dill
        with open(fileName, 'rb') as file:
            self.interpFn = dill.load(file)
        self.hasInterpIntegral = True

    def saveInterpolatingDensity(self, fileName):
        """ref: integrator.py:272-278"""
        import dill
        assert self.hasInterpDensity, \
            "No interpolating function to save"
        with open(fileName, 'wb') as file:
            dill.dump(self.interpDensity, file)

    def loadInterpolatingDensity(self, fileName):
        """ref: integrator.py:280-285"""
        import dill
        with open(fileName, 'rb') as file:
            self.interpDensity = dill.load(file)
        self.hasInterpDensity = True

    # ------------------
    # Analytic results: hooks for subclasses (ref: integrator.py:290-315)
    # ------------------
    def findAnalyticIntegral(self):
        """No analytic integral by default."""
        self.hasAnalyticIntegral = False

    def analyticIntegral(self, obs):
        """Analytic integral at the observable ``obs``, if a subclass knows one."""
        if not self.hasAnalyticIntegral:
            return
        pass

    def setAnalyticDensity(self, binspacing):
        """analyticDensity(observable): finite-difference derivative of the analytic integral at the bin midpoints."""
        from ..utils.hist_utils import histDerivative
        if not self.hasAnalyticIntegral:
            return
        self.analyticDensity = histDerivative(self.analyticIntegral(self.bin_midpoints), self.bins, giveHist=False,
                                              binInput=binspacing)
        self.hasAnalyticDensity = True

    # ------------------
    # csv save / load    (three files: values, errors, bin edges; ref: integrator.py:320-354)
    # ------------------
    @staticmethod
    def _write_csv(files, arrays):
        for name, arr in zip(files, arrays):
            np.savetxt(name, arr, delimiter=',')

    @staticmethod
    def _read_csv(files):
        return [np.loadtxt(name, delimiter=',') for name in files]

    def saveIntegral(self, intFile, errFile, binFile):
        assert self.hasMCIntegral, \
            "No integral histogram to save"
        self._write_csv((intFile, errFile, binFile), (self.integral, self.integralErr, self.bins))

    def loadIntegral(self, intFile, errFile, binFile):
        self.integral, self.integralErr, self.bins = self._read_csv((intFile, errFile, binFile))
        self.hasMCIntegral = True

    def saveDensity(self, denFile, errFile, binFile):
        assert self.hasMCDensity, \
            "No density histogram to save"
        self._write_csv((denFile, errFile, binFile), (self.density, self.densityErr, self.bins))

    def loadDensity(self, denFile, errFile, binFile):
        self.density, self.densityErr, self.bins = self._read_csv((denFile, errFile, binFile))
        self.hasMCDensity = True

    def checkValidAttributes(self):
        pass

    # state flags of the reference (integrator.py:367-386): nothing known yet
    _FLAGS = ('hasBins', 'hasMCDensity', 'hasMCIntegral', 'hasInterpIntegral', 'hasInterpDensity',
              'useFirstBinBndCond', 'useLastBinBndCond', 'monotone', 'hasAnalyticIntegral', 'hasAnalyticDensity')

    def __init__(self, *args, **kwargs):
        for flag in self._FLAGS:
            setattr(self, flag, False)
        self.firstBinBndCond = self.lastBinBndCond = None
        self.checkValidAttributes()


def integrate_1d(function, bounds, bin_space='lin', epsilon=1e-10, num_samples=1e5, num_bins=2,
                 bnd_cond=0, bnd_cond_bin='first'):
    """1-D Monte Carlo integral of ``function`` over ``bounds``.
    ref: montecarlo/integrator.py:390-457.  ``function`` is the caller's
    Python callable, so it runs on the host samples; sampling, histogramming
    and the cumulative sum run on the GPU."""
    if bin_space == 'lin':
        epsilon = None
    this_sampler = simpleSampler(bin_space, epsilon=epsilon, bounds=bounds)
    this_sampler.generateSamples(int(num_samples))
    samples = this_sampler.getSamples()

    this_integrator = integrator()
    if bnd_cond_bin == 'first':
        this_integrator.setFirstBinBndCondition(bnd_cond)
    elif bnd_cond_bin == 'last':
        this_integrator.setLastBinBndCondition(bnd_cond)
    else:
        raise AssertionError("Bin at which we place a boundary condition"
                             + "must be 'first' or 'last'.")

    this_integrator.setBins(num_bins, samples, bin_space)
    weights = function(samples)
    jacs = np.asarray(this_sampler.jacobians)
    area = this_sampler.area
    this_integrator.setDensity(samples, weights * jacs, area)
    this_integrator.integrate()

    integral = this_integrator.integral
    error = this_integrator.integralErr
    xs = this_integrator.bins
    if num_bins == 2:
        return integral[1], error[0], xs[1]
    return integral, error, xs


#######################################
# 2 dimensional integrator
#######################################
def multidim_cumsum(a):
    """Cumulative sum along every axis in turn, last axis first.  (Host helper kept for scripts that call it;
    integrator_2d does this on the device, jmc_finalize2d.)  ref: montecarlo/integrator.py:463-467"""
    out = a.cumsum(-1)
    for i in range(2, a.ndim + 1):
        np.cumsum(out, axis=-i, out=out)
    return out


class integrator_2d():
    """2-D version of ``integrator`` (ref: montecarlo/integrator.py:469-778; used by gen_pre_num_rad).
    np.histogram2d -> jmc_hist2d, normalisation and the double cumulative sum -> jmc_finalize2d.  The reference's
    quirks are kept: 'lin' bins of both axes start at min(0, min over BOTH observables) (:494-495), log
    bin_midpoints end up holding the y axis only (:508-511), density and integral are stored transposed
    ([y bin, x bin], :577-578)."""

    def _min_max(self, arr):
        dev = D.to_device(arr).reshape(-1)
        out = D.empty((2,))
        _lib.check(_lib.lib.jmc_minmax(D.ptr(dev), dev.numel(), D.ptr(out), D.stream()))
        return D.to_host(out)

    def setBins(self, numbins, observables, binspacing, min_log_bin=MIN_LOG_BIN):
        rng = [self._min_max(observables[i]) for i in range(2)]
        both_min = min(rng[0][0], rng[1][0]) if not (np.isnan(rng[0][0]) or np.isnan(rng[1][0])) else np.nan
        bins = []
        for i in range(2):
            if binspacing == 'lin':
                bins.append(np.linspace(min(0, both_min), rng[i][1], numbins))
            elif binspacing == 'log':
                min_bin = np.log10(max(rng[i][0], min_log_bin))
                max_bin = np.log10(rng[i][1])
                bins.append(np.logspace(min_bin, max_bin, numbins))
        self.bins = bins
        self.bin_midpoints = [[], []]
        xs, ys = bins[0], bins[1]
        if binspacing == 'lin':
            self.bin_midpoints[0] = (xs[1:] + xs[:-1]) / 2
            self.bin_midpoints[1] = (ys[1:] + ys[:-1]) / 2
        elif binspacing == 'log':
            self.bin_midpoints = np.exp((np.log(xs[1:]) + np.log(xs[:-1])) / 2.)
            self.bin_midpoints = np.exp((np.log(ys[1:]) + np.log(ys[:-1])) / 2.)
        self.hasBins = True

    def setLastBinBndCondition(self, bndCond):
        self.lastBinBndCond = bndCond
        self.useLastBinBndCond = True

    def setFirstBinBndCondition(self, bndCond):
        self.firstBinBndCond = bndCond
        self.useFirstBinBndCond = True

    def _finalize(self, area, n_total, bc_kind, bc_value, with_integral):
        nbx, nby = len(self.bins[0]) - 1, len(self.bins[1]) - 1
        dens, err = D.empty((nby, nbx)), D.empty((nby, nbx))
        integ = D.empty((nby, nbx)) if with_integral else None
        ierr = D.empty((nby, nbx)) if with_integral else None
        _lib.check(_lib.lib.jmc_finalize2d(
            D.ptr(self._sum_w), D.ptr(self._sum_w2), D.ptr(self._count), D.ptr(self._ex), nbx + 1, D.ptr(self._ey),
            nby + 1, float(area), int(n_total), bc_kind, float(bc_value), D.ptr(dens), D.ptr(err), D.ptr(integ),
            D.ptr(ierr), D.stream()))
        return dens, err, integ, ierr

    def setDensity(self, observables, weights, area):
        assert self.hasBins, "Need bins to produce density histogram."
        x = D.to_device(observables[0]).reshape(-1)
        y = D.to_device(observables[1]).reshape(-1)
        w = D.to_device(weights).reshape(-1)
        self._ex = D.to_device(np.asarray(self.bins[0], dtype=np.float64))
        self._ey = D.to_device(np.asarray(self.bins[1], dtype=np.float64))
        nbx, nby = self._ex.numel() - 1, self._ey.numel() - 1
        self._sum_w, self._sum_w2 = D.zeros((nbx, nby)), D.zeros((nbx, nby))
        self._count = D.zeros((nbx, nby), dtype=torch.int64)
        _lib.check(_lib.lib.jmc_hist2d(D.ptr(x), D.ptr(y), D.ptr(w), x.numel(), D.ptr(self._ex), nbx + 1,
                                       D.ptr(self._ey), nby + 1, D.ptr(self._sum_w), D.ptr(self._sum_w2),
                                       D.ptr(self._count), D.stream()))
        self._area, self._n_total = area, x.numel()
        dens, err, _, _ = self._finalize(area, x.numel(), _lib.BC_FIRST, 0., False)
        self.density, self.densityErr = D.to_host(dens), D.to_host(err)
        self.hasMCDensity = True

    def integrate(self):
        assert self.hasMCDensity, \
            "Need density function to perform integration"
        bc_kind, bc_value = _bc_args(self.firstBinBndCond, self.lastBinBndCond)
        self.useLastBinBndCond = bc_kind != _lib.BC_FIRST
        self.useFirstBinBndCond = bc_kind == _lib.BC_FIRST
        _, _, integ, ierr = self._finalize(self._area, self._n_total, bc_kind, bc_value, True)
        self.integral, self.integralErr = D.to_host(integ), D.to_host(ierr)
        self.hasMCIntegral = True
        self.set_extended_integral()

    def set_extended_integral(self):
        """integral with its boundary values appended (ref: integrator.py:650-685); host array assembly"""
        assert self.hasMCIntegral, "Cannot set extension of"\
            + " the integration to the boundaries without an integral."
        nx = len(self.bins[0])
        if self.useFirstBinBndCond:
            bc = self.firstBinBndCond
            z = [np.ones(nx) * bc] + [np.append(bc, z_i) for z_i in self.integral]
            zerr = [np.zeros(nx)] + [np.append(0, z_i) for z_i in self.integralErr]
        else:
            bc = self.lastBinBndCond[0]
            z = [np.append(z_i, bc) for z_i in self.integral] + [np.ones(nx) * bc]
            zerr = [np.append(0, z_i) for z_i in self.integralErr] + [np.zeros(nx)]
        self.extended_integral = np.array(z).T
        self.extended_integralErr = np.array(zerr).T

    def montecarlo_data_dict(self, info=None):
        data_dict = {'bins': self.bins,
                     'density': self.density,
                     'density error': self.densityErr,
                     'integral': self.extended_integral,
                     'integral error': self.extended_integralErr}
        if info is not None:
            data_dict['info'] = info
        return data_dict

    def save_montecarlo_data(self, filename, info=None):
        assert self.hasMCDensity and self.hasMCIntegral, \
            "Need density function and integral to save MC data"
        np.savez(filename, **self.montecarlo_data_dict(info=info))

    def makeInterpolatingFn(self, **kwargs):
        from ..utils.interpolation import get_2d_interpolation
        assert self.hasMCIntegral, \
            "Need MC integral to produce interpolation"
        self.interpFn = get_2d_interpolation(self.bins[0], self.bins[1], self.extended_integral, **kwargs)
        self.interpErr = None
        self.hasInterpIntegral = True

    def __init__(self, *args, **kwargs):
        for flag in integrator._FLAGS:
            if flag != 'monotone':
                setattr(self, flag, False)
        self.firstBinBndCond = self.lastBinBndCond = None


def integrate_2d(function, bounds, bin_space='lin', epsilon=1e-10, num_samples=1e5, num_bins=2, bnd_cond=0,
                 bnd_cond_bin='first'):
    """2-D Monte Carlo integral over [[xmin, xmax], [ymin, ymax]].  ref: montecarlo/integrator.py:780-858"""
    if bin_space == 'lin':
        epsilon = None
    s1 = simpleSampler(bin_space, bounds=bounds[0], epsilon=epsilon)
    s1.generateSamples(int(num_samples))
    s2 = simpleSampler(bin_space, bounds=bounds[1], epsilon=epsilon)
    s2.generateSamples(int(num_samples))
    x, y = s1.getSamples(), s2.getSamples()
    it = integrator_2d()
    if bnd_cond_bin == 'first':
        it.setFirstBinBndCondition(bnd_cond)
    elif bnd_cond_bin == 'last':
        it.setLastBinBndCondition(bnd_cond)
    else:
        raise AssertionError("Bin at which we place a boundary condition"
                             + "must be 'first' or 'last'.")
    it.setBins(num_bins, [x, y], bin_space)
    weights = function(x, y)
    it.setDensity([x, y], weights, s1.area * s2.area)
    it.integrate()
    integral, error = it.integral, it.integralErr
    if bnd_cond_bin == 'first':
        xs, ys = it.bins[0][1:], it.bins[1][1:]
    else:
        xs, ys = it.bins[0][:-1], it.bins[1][:-1]
    if num_bins == 2:
        return integral[0][0], error[0][0], xs[0], ys[0]
    return integral, error, xs, ys
