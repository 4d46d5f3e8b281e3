"""e+ e- -> hadrons at leading order -- mirror of the reference's one MonteCarloProcess client,
examples/processes/ee2hadrons_LO.py (process :13-71, pairwise observable :75-92, observable functions :95-150).

The process' phase-space constraint and differential cross section exist as device code
(JMC_PROCESS_EE2HADRONS_LO, csrc/jmc_process.cu); the Python callbacks below give the same values for single points.
The pairwise observable functions carry the id of their device implementation (``pair_id``), which
``ee2hLO_PairwiseObservable`` and ``montecarlo.ewoc.fused_ewoc`` use; called directly they work on NumPy arrays.
"""
from dataclasses import dataclass

import numpy as np

from .. import _device as D
from .. import _lib
from ..analytics.ewocs.eec import ee2hadrons_LO_prefactor
from ..montecarlo.process import MonteCarloObservable, MonteCarloProcess

import ctypes as C


class ee2hadrons_LO(MonteCarloProcess):
    r"""A Monte Carlo process for $e^+ e^- \to$ hadrons (at LO).  ref: examples/processes/ee2hadrons_LO.py:13-71"""
    device_process = _lib.PROCESS_EE2HADRONS_LO

    def _phasespace_constraints(self, **kwargs):
        """0 < x_i < 1 for the three energy fractions"""
        x_1, x_2 = 1 - kwargs['1 - x_1'], 1 - kwargs['1 - x_2']
        x_3 = 2 - x_1 - x_2
        return 0 < x_1 < 1 and 0 < x_2 < 1 and 0 < x_3 < 1

    def _differential_xsec(self, **kwargs):
        """The reference's expression as written: its x_2 is the kinematic variable ``1 - x_2`` itself (:28)."""
        prefactor = ee2hadrons_LO_prefactor(self.energy)
        x_1, x_2 = 1 - kwargs['1 - x_1'], kwargs['1 - x_2']
        return prefactor * (x_1**2. + x_2**2.) / ((1. - x_1) * (1. - x_2))

    def _device_prefactor(self):
        return ee2hadrons_LO_prefactor(self.energy)

    def named_samples(self):
        """x_1, x_2, x_3 of every sample"""
        samples = np.asarray(self.samples).reshape(-1, 2)
        x_1 = list(1 - samples[:, 0])
        x_2 = list(1 - samples[:, 1])
        x_3 = [2 - x_1[i] - x_2[i] for i in range(len(x_1))]
        return {'x_1': x_1, 'x_2': x_2, 'x_3': x_3}

    def __init__(self, energy, num_samples, sampleMethod='log', seed=None, **kwargs):
        kwargs.update({
            'name': r'$e^+ e^- \to$ hadrons (LO)',
            'accuracy': 'LO',
            'num_samples': num_samples,
            'sampleMethod': sampleMethod,
            'epsilon': 1e-8,
            'energy': energy,
            'kinematic_vars': ['1 - x_1', '1 - x_2'],
            'kinematic_bounds': [(0, 1), (0, 1)],
            # (the reference's placeholder: total cross section 1)
            'total_cross_section': 1.,
        })
        super().__init__(**kwargs)
        self.generateSamples(num_samples, seed=seed)
        self.setArea()


@dataclass
class ee2hLO_PairwiseObservable(MonteCarloObservable):
    """A pairwise observable over the three parton pairs (1,2), (2,3), (3,1) of every event, concatenated pair by
    pair; the weights are the event weights three times over.  ref: examples/processes/ee2hadrons_LO.py:75-92"""

    def update(self):
        p = self.process
        pair_id = getattr(self.kinematic_function, 'pair_id', None)
        self.device_observables = self.device_weights = None
        if pair_id is not None and getattr(p, 'device_samples', None) is not None \
                and p.device_samples.shape[0] == len(p.samples):
            n = p.device_samples.shape[0]
            out = D.empty((3 * n,))
            desc = p._descriptor()
            _lib.check(_lib.lib.jmc_process_pairs(C.byref(desc), D.ptr(p.device_samples), n, pair_id,
                                                  float(getattr(self.kinematic_function, 'kappa', 0.)), D.ptr(out),
                                                  D.stream()))
            self.device_observables = out
            self.device_weights = D.empty((3 * n,))
            _lib.check(_lib.lib.jmc_tile(D.ptr(p.device_jacobians), n, 3, D.ptr(self.device_weights), D.stream()))
            self.observables = D.to_host(out)
            self.weights = np.tile(np.asarray(p.jacobians, dtype=float), 3)
            return
        named = p.named_samples()
        x_1, x_2 = np.array(named['x_1']), np.array(named['x_2'])
        x_3 = 2 - x_1 - x_2
        self.observables = np.concatenate([self.kinematic_function(x_i, x_j)
                                           for (x_i, x_j) in [(x_1, x_2), (x_2, x_3), (x_3, x_1)]])
        self.weights = np.tile(np.asarray(p.jacobians, dtype=float), 3)


def _device(pair_id, **attrs):
    def mark(fn):
        fn.pair_id = pair_id
        for k, v in attrs.items():
            setattr(fn, k, v)
        return fn
    return mark


# ---- the observable functions (examples/processes/ee2hadrons_LO.py:95-150) --------------------------------------
@_device(_lib.PAIR_COSTHETA)
def costheta_ij(x_i, x_j):
    """cosine of the opening angle of partons i and j"""
    return 2. / (x_i * x_j) - 2. / x_i - 2. / x_j + 1


@_device(_lib.PAIR_THETA)
def theta_ij(x_i, x_j):
    return np.arccos(costheta_ij(x_i, x_j))


@_device(_lib.PAIR_M2)
def m2_ij(x_i, x_j):
    """pair mass squared over Q^2"""
    return -(1 - x_i - x_j)


def gecf_ij(x_i, x_j, kappa):
    """generalised energy correlation function of the pair"""
    return 2**(kappa / 2 - 2) * x_i * x_j *\
        (1 - costheta_ij(x_i, x_j))**(kappa / 2)


def gecf(kappa):
    """gecf_ij at fixed kappa as a two-argument pairwise observable (with its device id)"""
    return _device(_lib.PAIR_GECF, kappa=float(kappa))(lambda x_i, x_j: gecf_ij(x_i, x_j, kappa))


@_device(_lib.PAIR_Z_I)
def z_i(x_i, x_j):
    """energy weight of particle i"""
    return x_i / 2


@_device(_lib.PAIR_Z_J)
def z_j(x_i, x_j):
    """energy weight of particle j"""
    return x_j / 2
