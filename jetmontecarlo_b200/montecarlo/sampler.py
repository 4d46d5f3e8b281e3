"""Phase-space samplers -- host mirror of jetmontecarlo/montecarlo/sampler.py.

Same constructor arguments, methods, attributes and assertion messages as the
reference.  What changes is ``generateSamples``: the reference runs a Python
loop calling ``random.random()`` once per coordinate (sampler.py:30-34); here
one CUDA kernel (jmc_sample) draws all samples from a counter-based Philox
stream and applies the lin/log inverse transform.  The samples are written to
HBM only because this API hands them to the caller; the fused path
(jetmontecarlo_b200.fused) never stores them.

Additive API: ``generateSamples(num, seed=None)``, ``set_seed(seed)``,
``device_samples`` / ``device_jacobians`` (CUDA tensors of the same data).
"""
import ctypes as C
from abc import ABC, abstractmethod

import numpy as np

from ..utils.montecarlo_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from .. import _device as D
from .. import _lib

from .._seed import _mix64, next_key, set_seed  # noqa: F401  (the global Philox key sequence)


class sampler(ABC):
    """Abstract Monte Carlo phase-space sampler (ref: montecarlo/sampler.py:8-108)."""

    _kind = None      # _lib.SAMPLER_*
    _dim = 1

    # ------------------
    # Sampling Methods:
    # ------------------
    @abstractmethod
    def setArea(self):
        """Sets the area of the sampled phase space."""

    def _descriptor(self):
        return _lib.Sampler(kind=self._kind, method=_lib.METHODS[self.sampleMethod],
                            epsilon=float(self.epsilon or 0.), zc=float(getattr(self, 'zc', 0.) or 0.),
                            radius=float(getattr(self, 'radius', 1.)),
                            lo=float(getattr(self, 'bounds', [0, 1])[0]),
                            hi=float(getattr(self, 'bounds', [0, 1])[1]))

    def _draw(self, num, seed, first_index=0):
        D.require_cuda()
        shape = (num, 2) if self._dim == 2 else (num,)
        samples = D.empty(shape)
        jac = D.empty((num,))
        desc = self._descriptor()
        _lib.check(_lib.lib.jmc_sample(C.byref(desc), num, seed, first_index, D.ptr(samples), D.ptr(jac),
                                       D.stream()))
        return samples, jac

    @staticmethod
    def _next_key():
        return next_key()

    def __getstate__(self):
        """Pickled as the reference's samplers are (file_management.save_new_data(sampler, ..., '.pkl'),
        examples/event_generation/phase_space_sampling.py:145-148): the host arrays carry the data, the device copies
        are dropped and re-uploaded on demand."""
        state = dict(self.__dict__)
        for key in ('device_samples', 'device_jacobians'):
            if key in state:
                state[key] = None
        return state

    def addSample(self):
        """Adds a single sample (and its jacobian) to the lists."""
        samples, jac = self._draw(1, next_key())
        self.samples.append(D.to_host(samples)[0].tolist() if self._dim == 2 else float(D.to_host(samples)[0]))
        self.jacobians = list(self.jacobians) + [float(D.to_host(jac)[0])]

    def generateSamples(self, num, seed=None):
        """Generates ``num`` samples on the GPU (ref: sampler.py:30-34).

        As in the reference, ``samples`` becomes an ndarray, so a second call
        without ``clearSamples`` fails, and ``jacobians`` keeps growing across
        calls (``clearSamples`` does not reset it)."""
        if isinstance(self.samples, np.ndarray):
            raise AttributeError("'numpy.ndarray' object has no attribute 'append'")
        num = int(num)
        self.seed = next_key() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        dev_samples, dev_jac = self._draw(num, self.seed)
        self.device_samples, self.device_jacobians = dev_samples, dev_jac
        new = D.to_host(dev_samples)
        if len(self.samples):
            new = np.concatenate([np.array(self.samples).reshape((-1,) + new.shape[1:]), new])
        self.samples = new
        # kept as an ndarray (the reference appends Python floats to a list; building a list of 1e7 floats would
        # cost more than drawing the samples): np.array(s.jacobians), weights * s.jacobians, len() work unchanged
        new_jac = D.to_host(dev_jac)
        self.jacobians = new_jac if len(self.jacobians) == 0 else np.concatenate([np.asarray(self.jacobians, float),
                                                                                   new_jac])

    def getSamples(self):
        return self.samples

    def getSampleMethod(self):
        return self.sampleMethod

    def clearSamples(self):
        self.samples = []

    # ------------------
    # Saving/Loading:       (npz layout of the reference: keys 'samples', 'area')
    # ------------------
    def saveSamples(self, filename):
        np.savez(filename, samples=self.samples, area=self.area)

    def loadSamples(self, filename):
        npzfile = np.load(filename, allow_pickle=True, mmap_mode='c')
        self.samples = npzfile['samples']
        self.area = npzfile['area']

    # ------------------
    # Validity Checks:
    # ------------------
    def checkSampleMethod(self):
        validSampleMethods = ['lin', 'log']
        assert self.sampleMethod in validSampleMethods, \
            str(self.sampleMethod) + "is not a supported sample method."

    def checkValidEpsilon(self):
        eps = self.epsilon
        eps = 0 if eps is None else eps
        valid_eps = 0 < eps < 1.
        if self.sampleMethod == 'log':
            assert valid_eps, \
                ("epsilon={epsilon} is invalid for {type} sampling."
                 .format(epsilon=eps, type=self.sampleMethod))

    @abstractmethod
    def __init__(self, sampleMethod, epsilon=0., **kwargs):
        self.sampleMethod = sampleMethod
        self.epsilon = epsilon
        self.samples = []
        self.jacobians = []
        self.device_samples = None
        self.device_jacobians = None
        self.seed = None
        self.checkSampleMethod()
        self.checkValidEpsilon()
        self.setArea()
        super().__init__(**kwargs)


class simpleSampler(sampler):
    """Samples one variable on ``bounds`` (ref: montecarlo/sampler.py:114-140)."""
    _kind = _lib.SAMPLER_SIMPLE
    _dim = 1

    def setArea(self):
        if self.sampleMethod == 'lin':
            self.area = self.bounds[1] - self.bounds[0]
        elif self.sampleMethod == 'log':
            self.area = np.log(1. / self.epsilon)

    def __init__(self, sampleMethod, bounds=[0, 1], **kwargs):
        self.bounds = bounds
        super().__init__(sampleMethod, **kwargs)
