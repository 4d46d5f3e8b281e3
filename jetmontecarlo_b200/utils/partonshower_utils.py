"""Angularity-ordered parton shower -- host mirror of the part of
jetmontecarlo/utils/partonshower_utils.py that is on the hot path:
``gen_jets`` (:294-328), ``getECFs_ungroomed`` (:445-499) and
``getECFs_softdrop`` (:676-790).

The reference materialises every jet as a tree of Parton objects with
3-vectors.  With ``split_soft=False`` (the only mode ``gen_jets`` uses) a jet
is fully described by its ordered (z, theta) chain, and a chain is a pure
function of (seed, jet index) through the counter RNG -- so ``gen_jets`` here
returns a light ``JetList`` *description* and the ECF extractors regenerate
the chains on the GPU (jmc_shower_dump), writing only the per-jet observable.
For pdfs of very many jets use ``jetmontecarlo_b200.fused.Integrand.shower``,
which bins on the fly.
"""
import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from ..analytics.qcd_utils import MU_NP, check_jet_type
from ..fused import PRECISIONS, Integrand
from ..montecarlo.sampler import next_key


class JetList:
    """``num_events`` showered jets, described rather than stored."""

    def __init__(self, num_events, beta, jet_type, radius, acc, cutoff, seed):
        self.num_events = int(num_events)
        self.beta, self.jet_type, self.radius, self.acc, self.cutoff = beta, jet_type, radius, acc, cutoff
        self.seed = seed

    def __len__(self):
        return self.num_events

    def _integrand(self, **obs):
        return Integrand.shower(beta=self.beta, jet_type=self.jet_type, acc=self.acc, radius=self.radius,
                                cutoff=self.cutoff, **obs)

    def _dump(self, integrand, precision='f64'):
        D.require_cuda()
        n = self.num_events
        obs = D.empty((n,))
        n_em = D.empty((n,), dtype=torch.int32)
        n_tr = D.empty((n,), dtype=torch.int32)
        s = integrand.c_struct()
        _lib.check(_lib.lib.jmc_shower_dump(C.byref(s), n, self.seed, 0, PRECISIONS[precision], D.ptr(obs),
                                            D.ptr(n_em), D.ptr(n_tr), D.stream()))
        return obs, n_em, n_tr

    def chains(self, max_emissions=64, precision='f64'):
        """The event record: (lengths [n], z [n, max_emissions], theta [n, max_emissions]) of the ordered emission
        chains (NaN beyond a jet's length).  This is everything the reference's Parton trees hold when
        split_soft=False; any per-jet observable can be formed from it."""
        D.require_cuda()
        n = self.num_events
        chains = torch.full((n, int(max_emissions), 2), float('nan'), dtype=torch.float64, device=D.device())
        lens = D.empty((n,), dtype=torch.int32)
        s = self._integrand().c_struct()
        _lib.check(_lib.lib.jmc_shower_chains(C.byref(s), n, self.seed, 0, PRECISIONS[precision], int(max_emissions),
                                              D.ptr(chains), D.ptr(lens), D.stream()))
        c = D.to_host(chains)
        return D.to_host(lens), c[:, :, 0], c[:, :, 1]

    def emission_counts(self):
        """(accepted emissions, veto trials) per jet"""
        _, n_em, n_tr = self._dump(self._integrand())
        return D.to_host(n_em), D.to_host(n_tr)


def gen_jets(num_events, beta=2., jet_type='quark', radius=1., acc='LL', cutoff=MU_NP, verbose=1, seed=None):
    """ref: utils/partonshower_utils.py:294-328"""
    check_jet_type(jet_type)
    assert acc in ['LL', 'MLL'], "Invalid accuracy."
    if verbose > 0:
        print("Generating {:.0e} shower events".format(num_events)
              + " ordered by e_1^{(" + str(beta) + ")} ...", flush=True)
    return JetList(num_events, beta, jet_type, radius, acc, cutoff, next_key() if seed is None else int(seed))


def _check_n_emissions(n_emissions):
    assert n_emissions == 'all' or (isinstance(n_emissions, (int, np.integer)) and 1 <= n_emissions <= 4), \
        "n_emissions must be 1..4 or 'all' on the accelerated path"


def getECFs_ungroomed(jet_list, beta=2., obs_acc='LL', n_emissions=1, verbose=0):
    """List of ungroomed ECFs, one per jet.  ref: utils/partonshower_utils.py:445-499"""
    assert obs_acc in ['LL', 'MLL'], "Accuracy must be LL or MLL"
    _check_n_emissions(n_emissions)
    ig = jet_list._integrand(obs_acc=obs_acc, n_emissions=n_emissions)
    ig.beta_obs = beta
    assert beta == jet_list.beta, "the accelerated path bins the shower's own angularity exponent"
    obs, _, _ = jet_list._dump(ig)
    return list(D.to_host(obs))


def getECFs_softdrop(jet_list, z_cut, beta=2., beta_sd=0., obs_acc='LL', n_emissions=1, verbose=0,
                     reproduce_approximations=True):
    """List of Soft Drop groomed ECFs, one per jet.  ref: utils/partonshower_utils.py:676-790
    (per-jet rule of the docstring; the reference decrements ``n_emissions`` in place at :765, so with
    n_emissions=2 only its first jet really uses two emissions)."""
    assert obs_acc in ['LL', 'MLL'], "Accuracy must be LL or MLL"
    assert reproduce_approximations, "only the analytic-approximation mode is on the accelerated path"
    _check_n_emissions(n_emissions)
    assert beta == jet_list.beta, "the accelerated path bins the shower's own angularity exponent"
    ig = jet_list._integrand(obs_acc=obs_acc, n_emissions=n_emissions, groomer='softdrop', z_cut=z_cut,
                             beta_sd=beta_sd)
    obs, _, _ = jet_list._dump(ig)
    return list(D.to_host(obs))
