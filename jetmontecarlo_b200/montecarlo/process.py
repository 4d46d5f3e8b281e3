"""Physical processes sampled by hit-or-miss Monte Carlo -- host mirror of jetmontecarlo/montecarlo/process.py
(BasicProcess :14-28, KinematicProcess :46-103, MonteCarloProcess :109-191, MonteCarloObservable :202-237): same
class names, constructor keywords, methods and attributes.

What changes is where the sampling loop runs.  The reference's ``addSample`` draws every kinematic variable in its
box with ``random.random()``, asks two Python callbacks (phase-space constraint, differential cross section) and
repeats until a point is accepted -- once per sample, in Python.  A callback is code, so on the device a process is
one of the kinds compiled into libjmc_b200.so (``device_process``: include/jmc_b200.h, JMC_PROCESS_*): all samples
are drawn by ONE kernel (jmc_process_sample), sample i being the first accepted trial of Philox stream i, and the
rejected trials are counted for the efficiency-scaled area.  A subclass without a device kind cannot be sampled here
(there is no CPU path in this package); its Python callbacks remain available for single points.
"""
import ctypes as C
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Callable

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from .sampler import next_key, sampler


# =====================================
# Basic Process
# =====================================
class BasicProcess():
    """Name, energy and accuracy of a physical process.  ref: montecarlo/process.py:14-28"""

    def __str__(self):
        return f"The process [{self.name}] "\
            f"at an energy of {self.energy} GeV."\

    def __init__(self, name, energy, accuracy=None, **kwargs):
        self.name = name
        self.energy = energy
        self.accuracy = accuracy


# =====================================
# Kinematic Process
# =====================================
def assert_kinematic_kwargs(func):
    """The keyword arguments of a phase-space function must be exactly the process' kinematic variables.
    ref: montecarlo/process.py:34-45"""
    def wrapper(*args, **kwargs):
        kinematic_vars = args[0].kinematic_vars
        assert set(kwargs.keys()) == set(kinematic_vars), \
            "The parameters given to the function are not"\
            "the same as the kinematic variables."
        return func(*args, **kwargs)
    return wrapper


class KinematicProcess(ABC, BasicProcess):
    """Kinematic variables, phase space and differential cross section of a process.
    ref: montecarlo/process.py:46-103"""

    @assert_kinematic_kwargs
    def lies_in_phasespace(self, **kwargs):
        return self._phasespace_constraints(**kwargs)

    @abstractmethod
    def _phasespace_constraints(self, **kwargs):
        """The phase-space constraints of the process (subclass)."""

    @assert_kinematic_kwargs
    def differential_xsec(self, **kwargs):
        return self._differential_xsec(**kwargs)

    @abstractmethod
    def _differential_xsec(self, **kwargs):
        """The differential cross section of the process (subclass)."""

    def __init__(self, kinematic_vars: list, total_cross_section=None, **kwargs):
        self.kinematic_vars = kinematic_vars
        self.total_cross_section = total_cross_section
        self.num_variables = len(self.kinematic_vars)
        super().__init__(**kwargs)


# =====================================
# Monte Carlo Process
# =====================================
class MonteCarloProcess(sampler, KinematicProcess):
    """A kinematic process that samples its own phase space.  ref: montecarlo/process.py:109-191"""

    device_process = None        # _lib.PROCESS_* of a subclass whose callbacks exist as device code

    # ---------------------------------
    # Monte Carlo Sampling
    # ---------------------------------
    def _descriptor(self):
        if self.device_process is None:
            raise NotImplementedError(
                "%s has no device implementation (device_process): its phase-space constraint and differential cross "
                "section are Python callbacks, and jetmontecarlo_b200 has no CPU sampling path." % type(self).__name__)
        assert len(self.kinematic_bounds) == 2, "the device processes have two kinematic variables"
        lo = (C.c_double * 4)(*[float(b[0]) for b in self.kinematic_bounds], 0., 0.)
        hi = (C.c_double * 4)(*[float(b[1]) for b in self.kinematic_bounds], 0., 0.)
        return _lib.Process(kind=self.device_process, method=_lib.METHODS[self.sampleMethod],
                            epsilon=float(self.epsilon or 0.), lo=lo, hi=hi, energy=float(self.energy),
                            prefactor=float(self._device_prefactor()))

    def _device_prefactor(self):
        """constant factor of the differential cross section, evaluated once on the host (subclass)"""
        return 1.

    def _draw(self, num, seed, first_index=0):
        """(samples [num, d], weights [num], rejected trials) of the device process"""
        D.require_cuda()
        desc = self._descriptor()
        samples = D.empty((num, self.num_variables))
        weights = D.empty((num,))
        rejected = D.zeros((1,), dtype=torch.int64)
        _lib.check(_lib.lib.jmc_process_sample(C.byref(desc), num, seed, first_index, D.ptr(samples), D.ptr(weights),
                                               D.ptr(rejected), D.stream()))
        return samples, weights, int(rejected.item())

    def addSample(self):
        """One accepted phase-space point (and the trials rejected on the way).  ref: montecarlo/process.py:117-143"""
        samples, weights, rejected = self._draw(1, next_key())
        self.samples.append(D.to_host(samples)[0].tolist())
        self.jacobians = list(self.jacobians) + [float(D.to_host(weights)[0])]
        self.num_rejected_samples += rejected

    def generateSamples(self, num, seed=None):
        """``num`` accepted points in one kernel; ``jacobians`` holds jacobian * differential cross section, as in the
        reference.  ref: montecarlo/sampler.py:30-34 over montecarlo/process.py:117-143"""
        if isinstance(self.samples, np.ndarray):
            raise AttributeError("'numpy.ndarray' object has no attribute 'append'")
        num = int(num)
        self.seed = next_key() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        dev_samples, dev_weights, rejected = self._draw(num, self.seed)
        self.device_samples, self.device_jacobians = dev_samples, dev_weights
        self.num_rejected_samples += rejected
        new = D.to_host(dev_samples)
        if len(self.samples):
            new = np.concatenate([np.array(self.samples).reshape((-1, self.num_variables)), new])
        self.samples = new
        new_w = D.to_host(dev_weights)
        self.jacobians = new_w if len(self.jacobians) == 0 else np.concatenate([np.asarray(self.jacobians, float), new_w])

    def named_samples(self):
        """{kinematic variable: its samples}.  ref: montecarlo/process.py:146-151"""
        return {varname: [sample[i] for sample in self.samples]
                for i, varname in enumerate(self.kinematic_vars)}

    def _max_area(self):
        """Volume of the box of kinematic bounds.  ref: montecarlo/process.py:154-161"""
        max_area = 1.
        for bound in self.kinematic_bounds:
            max_area *= bound[1] - bound[0]
        return max_area

    def setArea(self):
        """Box volume times the acceptance of the samples drawn so far.  ref: montecarlo/process.py:164-175"""
        if len(self.samples) == 0:
            self.area = None
            return
        efficiency = len(self.samples) \
            / (len(self.samples) + self.num_rejected_samples)
        self.area = self._max_area() * efficiency
        return

    def __init__(self, kinematic_bounds, **kwargs):
        self.kinematic_bounds = kinematic_bounds
        self.num_rejected_samples = 0
        super().__init__(**kwargs)
        assert len(self.kinematic_bounds) == self.num_variables, \
            "The number of bounds for the kinematic variables doesn't"\
            " match the number of kinematic variables."


# =====================================
# Observable for a Process
# =====================================
@dataclass
class MonteCarloObservable:
    """An observable of a MonteCarloProcess: a function of its kinematic variables, evaluated on its samples, with
    the process' weights.  ref: montecarlo/process.py:202-237"""
    process: MonteCarloProcess
    kinematic_function: Callable
    name: str = ''

    def update(self):
        samples = np.asarray(self.process.samples)
        try:        # array-capable functions in one call; anything else sample by sample as the reference does
            self.observables = list(np.asarray(self.kinematic_function(*samples.T), dtype=float))
        except Exception:
            self.observables = [self.kinematic_function(*sample) for sample in self.process.samples]
        self.weights = self.process.jacobians

    def __str__(self):
        return f"Observables {self.name} for process {self.process.name}:\n"\
            + f"{self.observables}\n"

    def __post_init__(self, *args, **kwargs):
        self.observables = None
        self.weights = None
        self.update()
