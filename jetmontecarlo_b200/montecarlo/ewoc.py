"""Energy-weighted observable correlators -- host mirror of jetmontecarlo/montecarlo/ewoc.py
(Npoint_MonteCarloEWOC :11-87, Npoint_MonteCarloSubJetEWOC :93-137; plot_ewoc is plotting and not mirrored).

The arrays stay on the device when the observables came from a device process: the energy weights
``z1**w1 * z2**w2 * weights`` are one kernel (JMC_FN_EWOC_WEIGHT), the bin range one reduction (jmc_minmax), the
histogram one pass (jmc_hist).  ``fused_ewoc`` does the whole chain -- generate, three pairs per event, observable,
weights, histogram -- in one kernel without storing anything per event (jmc_process_ewoc).
"""
import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from .._arrayfn import eval_fn
from .integrator import integrator
from .process import MonteCarloObservable, MonteCarloProcess


def _values(observable, attr):
    """device tensor of an observable's values if it has one, else its host values"""
    dev = getattr(observable, 'device_' + attr, None)
    return dev if dev is not None else np.asarray(getattr(observable, attr), dtype=np.float64)


class Npoint_MonteCarloEWOC(integrator):
    """N-point (N = 2) energy-weighted correlator of a process observable.  ref: montecarlo/ewoc.py:11-87"""

    def set_weights(self, z1Observable: MonteCarloObservable, z2Observable: MonteCarloObservable,
                    constraintObservable: MonteCarloObservable = None):
        """weights <- z1**w1 * z2**w2 * weights.  ref: montecarlo/ewoc.py:14-28"""
        (w1, w2) = self.energy_weights
        self._set_energy_weights(_values(z1Observable, 'observables'), _values(z2Observable, 'observables'), w1, w2)

    def _set_energy_weights(self, z1, z2, w1, w2):
        out = eval_fn(_lib.FN_EWOC_WEIGHT, [z1, z2, self._weights_any, float(w1)], beta=float(w2))
        self._weights_any = out
        self.weights = D.to_host(out) if isinstance(out, torch.Tensor) else out
        self.has_weights = True

    def generate_ewoc(self, numbins, binspacing, minval=None, maxval=None):
        """Bins from the range of the observables (optionally widened to minval / maxval), then the weighted density
        with the process' area.  ref: montecarlo/ewoc.py:30-57"""
        if not self.has_weights:
            raise ValueError("Weights have not been set. Please use\n"
                             "    `ewocInstance.set_weights(z1_function, "
                             "z2_function)`,\n"
                             "where zi_function is the functional form "
                             "of the ith energy fraction in terms of "
                             "the kinematic variables of the process.")
        lo, hi = self._min_max(self._observables_any)
        # joining minval / maxval to the observables only moves the ends of the range setBins looks at
        if minval is not None:
            lo = min(minval, lo)
        if maxval is not None:
            hi = max(maxval, hi)
        self.setBins(numbins=numbins, observables=np.array([lo, hi]), binspacing=binspacing)
        self.setDensity(observables=self._observables_any, weights=self._weights_any,
                        area=self.montecarlo_process.area)

    def __init__(self, montecarlo_process: MonteCarloProcess, montecarlo_observable: MonteCarloObservable,
                 N=2, energy_weights=(1, 1)):
        self.montecarlo_process = montecarlo_process
        self.montecarlo_observable = montecarlo_observable
        self._weights_any = _values(montecarlo_observable, 'weights')
        self._observables_any = _values(montecarlo_observable, 'observables')
        self.weights = np.array(self.montecarlo_observable.weights)
        self.observables = np.array(self.montecarlo_observable.observables)
        self.N = N
        self.energy_weights = energy_weights
        self.has_weights = False
        assert self.N == 2, "Only N=2 is implemented"
        assert len(self.energy_weights) == self.N,\
            "Weights must be of length N"
        super().__init__()


class Npoint_MonteCarloSubJetEWOC(Npoint_MonteCarloEWOC):
    """The same for pairs inside a jet but outside a subjet (deprecated in the reference in favour of phase-space
    constraints on the process; kept for its scripts).  ref: montecarlo/ewoc.py:93-137"""

    def set_weights(self, z1Observable: MonteCarloObservable, z2Observable: MonteCarloObservable,
                    angleObservable: MonteCarloObservable):
        angles = np.asarray(angleObservable.observables)
        in_jet = self.jet_algorithm.is_in_jet(angles)
        in_subjet = self.subjet_algorithm.is_in_jet(angles)
        ewoc_inds = np.where(in_jet * np.logical_not(in_subjet))[0]
        # the selection changes the number of entries, which is the N of the density's normalisation
        self.observables = self.observables[ewoc_inds]
        self._observables_any = self.observables
        self._weights_any = np.asarray(self.weights)[ewoc_inds]
        (w1, w2) = self.energy_weights
        self._set_energy_weights(np.asarray(z1Observable.observables)[ewoc_inds],
                                 np.asarray(z2Observable.observables)[ewoc_inds], w1, w2)

    def __init__(self, jetAlgorithm, subjetAlgorithm, **kwargs):
        self.jet_algorithm = jetAlgorithm
        self.subjet_algorithm = subjetAlgorithm
        super().__init__(**kwargs)


def fused_ewoc(process, num_samples, pair_function, energy_weights=(1, 1), numbins=100, binspacing='lin',
               minval=None, maxval=None, edges=None, seed=None, kappa=0.):
    """The whole EWOC of a device process in one kernel per pass, nothing stored per event: a min/max pass for the
    bin range (skipped when ``edges`` are given), then generate -> three parton pairs per event -> observable,
    z_i^w1 z_j^w2 weight -> histogram (jmc_process_ewoc).  ``pair_function``: one of the pairwise observables of
    ``jetmontecarlo_b200.processes.ee2hadrons_LO`` (it carries its device id).  Returns an ``integrator`` with bins,
    density and densityErr; the process' area and rejected count are those of this run (``it.area``,
    ``it.num_rejected``).  ref: montecarlo/ewoc.py:11-57 over montecarlo/process.py:109-175"""
    from .sampler import next_key
    D.require_cuda()
    obs_id = getattr(pair_function, 'pair_id', None)
    assert obs_id is not None, "fused_ewoc needs a pairwise observable with a device implementation"
    kappa = float(getattr(pair_function, 'kappa', kappa))
    desc = process._descriptor()
    n, key = int(num_samples), (next_key() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF)
    (w1, w2) = energy_weights
    it = integrator()
    rejected = D.zeros((1,), dtype=torch.int64)
    if edges is None:
        mm = torch.tensor([np.inf, -np.inf], dtype=torch.float64, device=D.device())
        _lib.check(_lib.lib.jmc_process_ewoc(C.byref(desc), n, key, 0, obs_id, kappa, float(w1), float(w2), None, 0,
                                             None, None, None, D.ptr(mm), D.ptr(rejected), D.stream()))
        lo, hi = D.to_host(mm)
        if minval is not None:
            lo = min(minval, lo)
        if maxval is not None:
            hi = max(maxval, hi)
        it._edges_from_range(numbins, lo, hi, binspacing, 1e-15)
        D.zero_(rejected)
    else:
        it.bins, it.binspacing, it.hasBins = np.asarray(edges, dtype=np.float64), binspacing, True
        it.bin_midpoints = ((it.bins[1:] + it.bins[:-1]) / 2. if binspacing == 'lin'
                            else np.exp((np.log(it.bins[1:]) + np.log(it.bins[:-1])) / 2.))
    it._alloc_hist()
    _lib.check(_lib.lib.jmc_process_ewoc(C.byref(desc), n, key, 0, obs_id, kappa, float(w1), float(w2),
                                         D.ptr(it._edges_dev), len(it.bins), D.ptr(it._sum_w), D.ptr(it._sum_w2),
                                         D.ptr(it._count), None, D.ptr(rejected), D.stream()))
    it.num_rejected = int(rejected.item())
    it.area = process._max_area() * (n / (n + it.num_rejected))
    it._density_from_hist(it.area, 3 * n)
    return it
