"""Inverse-transform "Sudakov samples", one sample per event from a cdf that depends on the event -- the stage between
the numerical radiators and the final pdf in the reference's production workflow
(examples/sudakov_comparisons/sudakov_utils.py:100-238: ``generate_critical_samples``,
``generate_precritical_samples``, ``generate_subsequent_samples``).

The reference loops over events in Python; every iteration calls ``samples_from_cdf(cdf_i, 1, domain_i,
force_monotone=True)`` (utils/montecarlo_utils.py:93-375), i.e. evaluates the radiator interpolant on ~1000 probe
points and walks a list of known cdf shapes.  Here the radiator is what it is in that workflow -- a table on a
rectangular grid (``TableRadiator``, the result of gen_pre_num_rad / gen_crit_sub_num_rad) -- and ONE kernel
(jmc_conditional_inverse_transform, csrc/jmc_sudakov.cu) gives every event a thread block that keeps the event's
probe grid and tabulated cdf in shared memory and walks the same list of shapes.  Conditions, uniforms and samples
are the only per-event HBM traffic.
"""
import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib


class TableRadiator:
    """A numerical radiator R(x, y) tabulated on a rectangular grid, resident on the device: what
    ``get_2d_interpolation(xs, ys, zs)`` (utils/interpolation.py:248-273) interpolates.  ``method``:
    'RectangularGrid' (bilinear, linearly extrapolated outside the grid: the reference's default) or 'Nearest'.
    Calling it evaluates the interpolant on host arrays (scipy, as the host mirror of ``get_2d_interpolation``
    does); the sampling kernels read the table itself."""

    def __init__(self, xs, ys, zs, method="RectangularGrid"):
        assert method in ("RectangularGrid", "Nearest"), "Unknown interpolation method " + str(method)
        self.xs = np.ascontiguousarray(xs, dtype=np.float64)
        self.ys = np.ascontiguousarray(ys, dtype=np.float64)
        self.zs = np.ascontiguousarray(zs, dtype=np.float64)
        assert self.zs.shape == (len(self.xs), len(self.ys)), "table must be [len(xs), len(ys)]"
        assert np.all(np.diff(self.xs) > 0) and np.all(np.diff(self.ys) > 0), "grid axes must increase strictly"
        self.method = method
        self._dev = None

    def device(self):
        if self._dev is None:
            self._dev = (D.to_device(self.xs), D.to_device(self.ys), D.to_device(self.zs.reshape(-1)))
        return self._dev

    def __call__(self, x, y):
        from ..utils.interpolation import get_2d_interpolation
        x, y = np.broadcast_arrays(np.asarray(x, dtype=float), np.asarray(y, dtype=float))
        if self.method == "Nearest":
            i = np.abs(x[..., None] - self.xs).argmin(axis=-1)
            j = np.abs(y[..., None] - self.ys).argmin(axis=-1)
            return self.zs[i, j]
        return get_2d_interpolation(self.xs, self.ys, self.zs)(np.stack([x, y], axis=-1))


def grid_interpolant(xs, ys, zs, device=None, method="RectangularGrid"):
    """the radiator table of a gen_pre_num_rad / gen_crit_sub_num_rad result as a ``TableRadiator``"""
    return TableRadiator(xs, ys, zs, method)


def conditional_samples_from_table(radiator, conds, domain_lo, domain_hi, uniforms, catch_turnpoint=False,
                                   force_monotone=False):
    """One sample per event i from the cdf ``x -> exp(-R(x, conds[i]))`` on [domain_lo[i], domain_hi[i]] with the
    uniform ``uniforms[i]``: per event ``samples_from_cdf(cdf_i, 1, domain_i, catch_turnpoint, force_monotone)`` on
    its tabulated route.  All arguments but the radiator are CUDA tensors [N]; returns (samples [N], weights [N] = 1).
    ValueError if some event's cdf is not monotone and none of the options resolves it.
    ref: utils/montecarlo_utils.py:142-318, 377-399"""
    D.require_cuda()
    assert isinstance(radiator, TableRadiator), "the conditional sampler needs the radiator's table (TableRadiator)"
    conds = D.to_device(conds).reshape(-1)
    lo, hi, r = (D.to_device(t).reshape(-1) for t in (domain_lo, domain_hi, uniforms))
    n = conds.numel()
    assert lo.numel() == hi.numel() == r.numel() == n
    gx, gy, table = radiator.device()
    out = D.empty((n,))
    rc = _lib.lib.jmc_conditional_inverse_transform(D.ptr(gx), gx.numel(), D.ptr(gy), gy.numel(), D.ptr(table),
                                                    int(radiator.method == "Nearest"), D.ptr(conds), D.ptr(lo), D.ptr(hi),
                                                    D.ptr(r), n, int(bool(catch_turnpoint)), int(bool(force_monotone)),
                                                    D.ptr(out), D.stream())
    if rc == -5:        # JMC_EDOMAIN: what the reference raises as ValueError
        raise ValueError(_lib.lib.jmc_last_error().decode("utf-8", "replace"))
    _lib.check(rc)
    return out, D.full((n,), 1.)


# ---- the three generators of the production workflow ---------------------------------------------------------------
def _generate(emission, radiator, theta_crits, param, seed, uniforms):
    """jmc_sudakov_generate: domains, uniforms, sampling, inf -> (0, weight 0) formatting and weights in ONE kernel"""
    D.require_cuda()
    assert isinstance(radiator, TableRadiator), "the conditional sampler needs the radiator's table (TableRadiator)"
    from ..montecarlo.sampler import next_key
    theta_crits = D.to_device(theta_crits).reshape(-1)
    n = theta_crits.numel()
    r = None if uniforms is None else D.to_device(uniforms).reshape(-1)
    assert r is None or r.numel() == n
    key = 0 if r is not None else (next_key() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF)
    gx, gy, table = radiator.device()
    samples, weights = D.empty((n,)), D.empty((n,))
    rc = _lib.lib.jmc_sudakov_generate(emission, D.ptr(gx), gx.numel(), D.ptr(gy), gy.numel(), D.ptr(table),
                                       int(radiator.method == "Nearest"), D.ptr(theta_crits), float(param), D.ptr(r), key,
                                       n, D.ptr(samples), D.ptr(weights), D.stream())
    if rc == -5:        # JMC_EDOMAIN: what the reference raises as ValueError
        raise ValueError(_lib.lib.jmc_last_error().decode("utf-8", "replace"))
    _lib.check(rc)
    return samples, weights


def generate_precritical_samples(pre_radiator, theta_crits, z_cut, seed=None, uniforms=None):
    """z_pre for every critical angle, from the cdf exp(-R_pre(z_pre, theta_crit)) on [0, z_cut] forced monotone;
    an infinite sample comes back as 0 with weight 0.  ``pre_radiator``: ``TableRadiator`` of a gen_pre_num_rad
    result; ``theta_crits``: array / CUDA tensor.  Returns (samples, weights) CUDA tensors.
    ref: examples/sudakov_comparisons/sudakov_utils.py:134-182, 117-120"""
    return _generate(_lib.SUDAKOV_PRECRITICAL, pre_radiator, theta_crits, z_cut, seed, uniforms)


def generate_subsequent_samples(sub_radiator, theta_crits, beta, seed=None, uniforms=None):
    """C_sub for every critical angle, from exp(-R_sub(C, theta_crit)) on [0, theta_crit^beta / 2] forced monotone;
    events with theta_crit^beta / 2 < 1e-10 go to the underflow value 1e-100.
    ref: examples/sudakov_comparisons/sudakov_utils.py:185-238"""
    return _generate(_lib.SUDAKOV_SUBSEQUENT, sub_radiator, theta_crits, beta, seed, uniforms)


def generate_critical_samples(crit_radiator, num_samples, seed=None):
    """theta_crit samples from exp(-R_crit(theta)) on [0, 1] (host-callable radiator, e.g. the integrator's
    interpolating function): ``samples_from_cdf`` with the draw on the device.  Returns NumPy arrays.
    ref: examples/sudakov_comparisons/sudakov_utils.py:100-131"""
    from ..utils.montecarlo_utils import samples_from_cdf
    samples, weights = samples_from_cdf(lambda theta: np.exp(-1. * crit_radiator(theta)), int(num_samples),
                                        domain=[0., 1.], backup_cdf=None, verbose=0, seed=seed)
    weights = np.where(np.isinf(samples), 0, weights)
    return np.where(np.isinf(samples), 0, samples), weights
