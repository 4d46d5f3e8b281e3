"""Jet phase-space samplers -- host mirror of
jetmontecarlo/numerics/radiators/samplers.py (criticalSampler,
precriticalSampler, ungroomedSampler): same arguments, areas, jacobians and
assertion messages; sampling itself is the jmc_sample CUDA kernel.
"""
import numpy as np

from ... import _lib
from ...utils.montecarlo_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ...montecarlo.sampler import sampler


def checkValidzcut(zc):
    """ref: numerics/radiators/samplers.py:12-18"""
    zc = 0 if zc is None else zc
    valid_zc = 0 < zc < 1. / 2.
    assert valid_zc, \
        ("zc={zcut} is an invalid grooming parameter."
         .format(zcut=zc))


class criticalSampler(sampler):
    """z in (zc, 1/2), theta in (0, radius); log sampling: jac = (z - zc) theta.
    ref: numerics/radiators/samplers.py:23-119"""
    _kind = _lib.SAMPLER_CRITICAL
    _dim = 2

    def setArea(self):
        if self.sampleMethod == 'lin':
            self.area = (1. / 2. - self.zc) * self.radius
        elif self.sampleMethod == 'log':
            self.area = np.log(1. / self.epsilon) ** 2.
        else:
            raise ValueError('Invalid sample method.')

    def rescaleEnergies(self, zs):
        """Dead in the reference too (ends in ``assert False``, samplers.py:64-93)."""
        assert False, "Incomplete rescaling!"

    def __init__(self, sampleMethod, zc, epsilon=0., radius=1.):
        self.zc = zc
        checkValidzcut(zc)
        assert radius > 0, "Jet radius must be greater than zero!"
        self.radius = radius
        super().__init__(sampleMethod, epsilon)


class precriticalSampler(sampler):
    """z_pre in (0, zc); log sampling: jac = z_pre.
    ref: numerics/radiators/samplers.py:125-180"""
    _kind = _lib.SAMPLER_PRECRITICAL
    _dim = 1

    def setArea(self):
        if self.sampleMethod == 'lin':
            self.area = self.zc
        elif self.sampleMethod == 'log':
            self.area = np.log(1. / self.epsilon)
        else:
            raise ValueError('Invalid sample method.')

    def __init__(self, sampleMethod, zc, epsilon=0.):
        self.zc = zc
        checkValidzcut(zc)
        super().__init__(sampleMethod, epsilon)


class ungroomedSampler(sampler):
    """z in (0, 1/2), theta in (0, radius); log sampling: jac = z theta.
    ref: numerics/radiators/samplers.py:189-302"""
    _kind = _lib.SAMPLER_UNGROOMED
    _dim = 2

    def setArea(self):
        if self.sampleMethod == 'lin':
            self.area = self.radius / 2.
        elif self.sampleMethod == 'log':
            self.area = np.log(1. / self.epsilon) ** 2.
        else:
            raise ValueError('Invalid sample method.')

    def rescaleEnergies(self, zs):
        assert False, "Incomplete rescaling!"

    def rescaleAngles(self, thetas):
        assert False, "Incomplete rescaling!"

    def __init__(self, sampleMethod, epsilon=0., radius=1.):
        assert radius > 0, "Jet radius must be greater than zero!"
        self.radius = radius
        super().__init__(sampleMethod, epsilon)
