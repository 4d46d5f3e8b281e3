"""Angularity-ordered parton shower -- host mirror of the part of
jetmontecarlo/utils/partonshower_utils.py that is on the hot path:
``gen_jets`` (:294-328), ``getECFs_ungroomed`` (:445-499),
``getECFs_softdrop`` (:676-790), ``getECFs_rss`` (:501-674), ``getangs``
(:391-443), the ``getECFs`` dispatcher and ``save_shower_correlations``
(:792-884, the production script's output file).

The reference materialises every jet as a tree of Parton objects with
3-vectors.  With ``split_soft=False`` (the only mode ``gen_jets`` uses) a jet
is fully described by its ordered (z, theta) chain, and a chain is a pure
function of (seed, jet index) through the counter RNG -- so ``gen_jets`` here
returns a light ``JetList`` *description* and the ECF extractors regenerate
the chains on the GPU (jmc_shower_dump), writing only the per-jet observable.
For pdfs of very many jets use ``jetmontecarlo_b200.fused.Integrand.shower``,
which bins on the fly.  RSS grooming runs inside the shower kernel as well
(per-lane state of the jet being generated); the event record
(``JetList.chains_device``) is written only on request.
"""
import ctypes as C

import numpy as np
import torch

from .. import _device as D
from .. import _lib
from ..analytics.qcd_utils import MU_NP, check_jet_type
from ..fused import PRECISIONS, Integrand
from ..montecarlo.sampler import next_key


CHAIN_RECORD = 64        # emissions per jet in a first event record (MLL gluon jets average about 20)


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

    def chains_device(self, max_emissions=64, precision='f64'):
        """The event record on the device: (lengths [n] int32, chains [n, max_emissions, 2] = (z, theta), NaN
        beyond a jet's length)."""
        D.require_cuda()
        n = self.num_events
        chains = torch.full((n, int(max_emissions), 2), float('nan'), dtype=torch.float64, device=D.device())
        lens = D.empty((n,), dtype=torch.int32)
        s = self._integrand().c_struct()
        _lib.check(_lib.lib.jmc_shower_chains(C.byref(s), n, self.seed, 0, PRECISIONS[precision], int(max_emissions),
                                              D.ptr(chains), D.ptr(lens), D.stream()))
        return lens, chains

    def chains(self, max_emissions=64, precision='f64'):
        """The event record: (lengths [n], z [n, max_emissions], theta [n, max_emissions]) of the ordered emission
        chains (NaN beyond a jet's length).  This is everything the reference's Parton trees hold when
        split_soft=False; any per-jet observable can be formed from it."""
        lens, chains = self.chains_device(max_emissions, precision)
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



def getECFs_rss(jet_list, z_cut, beta=2., f=1., obs_acc='LL', emission_type='precritsub', n_emissions=1,
                few_pres=True, rescale_hard=True, verbose=0):
    """List of RSS-groomed ECFs, one per jet, from *ungroomed* jets.  Recursive safe subtraction runs inside the shower
    kernel, on the per-lane state of the jet being generated (``EcfAcc::add_rss``, csrc/jmc_shower.cu): emissions
    softer than what is left of z_cut are pre-critical (they use z_cut up and, for f != 1, rescale the hard branch),
    the first survivor is critical, later ones are subsequent; the ECF is the sum of the ``n_emissions`` largest
    contributions (or of all).  ``emission_type`` is matched by substring, as the reference does.  Only the per-jet
    observable comes back.  ref: utils/partonshower_utils.py:501-674"""
    if z_cut == 0:
        return getECFs_ungroomed(jet_list, beta=beta, obs_acc=obs_acc, n_emissions=n_emissions, verbose=verbose)
    assert obs_acc in ['LL', 'MLL'], "Accuracy must be LL or MLL"
    assert f > 1 / 3, "f = " + str(f) + " is less than 1/3, and the grooming may groom away harder branches first."
    _check_n_emissions(n_emissions)
    assert beta == jet_list.beta, "the accelerated path bins the shower's own angularity exponent"
    ig = jet_list._integrand(obs_acc=obs_acc, n_emissions=n_emissions, groomer='rss', z_cut=z_cut, f=f,
                             emission_type=emission_type, few_pres=few_pres)
    obs, _, _ = jet_list._dump(ig)
    ecfs = D.to_host(obs)
    return list(ecfs[ecfs >= 0])                # the reference keeps a jet only if its ecf >= 0


def ecfs_from_chains(lens, chains, beta=2., obs_acc='LL', n_emissions=1, groomer='ungroomed', z_cut=0., beta_sd=0.,
                     f=1., emission_type='precritsub', few_pres=True):
    """The ECF pickers on a stored event record (``JetList.chains_device`` or any (z, theta) chains): CUDA tensors
    ``lens`` [n] int32 and ``chains`` [n, L, 2] -> tensor [n] of observables, evaluated by ``jmc_chain_ecf`` with the
    accumulator of the shower kernel (FP64, the reference's operation order).
    ref: utils/partonshower_utils.py:445-499, 501-674, 676-790"""
    D.require_cuda()
    _check_n_emissions(n_emissions)
    ig = Integrand.shower(beta=beta, obs_acc=obs_acc, n_emissions=n_emissions, groomer=groomer, z_cut=z_cut,
                          beta_sd=beta_sd, f=f, emission_type=emission_type, few_pres=few_pres)
    lens = lens.to(device=D.device(), dtype=torch.int32).contiguous()
    chains = chains.to(device=D.device(), dtype=torch.float64).contiguous()
    assert chains.dim() == 3 and chains.shape[0] == lens.numel() and chains.shape[2] == 2
    assert lens.numel() == 0 or int(lens.max().item()) <= chains.shape[1], \
        "emission chains were truncated: raise max_emissions"
    out = D.empty((lens.numel(),))
    s = ig.c_struct()
    _lib.check(_lib.lib.jmc_chain_ecf(C.byref(s), D.ptr(lens), D.ptr(chains), lens.numel(), int(chains.shape[1]),
                                      D.ptr(out), D.stream()))
    return out


def getangs(jet_list, beta=2., acc='LL', emission_type=None, verbose=1):
    """List of angularities max_i z_i theta_i^beta, one per jet (1e-50 for a jet without emissions).
    ref: utils/partonshower_utils.py:391-443 -- at 'LL' and 'MLL' alike the reference takes the largest single
    emission of z theta^beta, which is the LL ungroomed ECF of the leading emission."""
    assert emission_type is None, "the event record carries no emission types (the reference's jets from gen_jets do not either)"
    assert acc in ['LL', 'MLL'], "the reference's 'ME' branch raises (sum of a float, :430); LL and MLL are supported"
    return getECFs_ungroomed(jet_list, beta=beta, obs_acc='LL', n_emissions=1, verbose=0)


def getECFs(jet_list, params, groomer=None):
    """ref: utils/partonshower_utils.py:792-800"""
    if groomer is None:
        return getECFs_ungroomed(jet_list, **params)
    elif groomer == 'RSS':
        return getECFs_rss(jet_list, **params)
    elif groomer in ['SoftDrop', 'SD', 'Soft Drop']:
        return getECFs_softdrop(jet_list, **params)
    else:
        raise ValueError('Invalid groomer type ' + str(groomer) + '.')


def save_shower_correlations(jet_list, file_path, z_cuts=[.05, .1, .2], beta=2., beta_sd=0., f_soft=1.,
                             obs_acc='LL', few_pres=True, fixed_coupling=True, verbose=1):
    """Writes the correlations file of the shower production script: same keys and array shapes as the
    reference (ungroomed (1, N); RSS crit / critsub / precrit (n_zcut, N); RSS precritsub / two / all and the
    Soft Drop ones (n_zcut, 1, N)).  ref: utils/partonshower_utils.py:802-884"""
    n_emissions_list = [1, 2, 'all']
    ungroomed = [[getECFs_ungroomed(jet_list, beta=beta, obs_acc=obs_acc, n_emissions=n, verbose=verbose)]
                 for n in n_emissions_list]
    rss = {k: [] for k in ('crit', 'critsub', 'precrit', 'precritsub', 'two', 'all')}
    softdrop = {k: [] for k in ('crit', 'two', 'all')}
    for z_cut in z_cuts:
        params = dict(jet_list=jet_list, z_cut=z_cut, beta=beta, f=f_soft, obs_acc=obs_acc, few_pres=few_pres,
                      verbose=verbose)
        for etype in ('crit', 'precrit', 'critsub'):
            rss[etype].append(getECFs_rss(**params, emission_type=etype))
        for key, n in zip(('precritsub', 'two', 'all'), n_emissions_list):
            rss[key].append([getECFs_rss(**params, emission_type='precritsub', n_emissions=n)])
    for z_cut in z_cuts:
        # (beta_sd = 0 whatever the argument says, as in the reference :853)
        params = dict(jet_list=jet_list, z_cut=z_cut, beta=beta, beta_sd=0., obs_acc=obs_acc, verbose=verbose)
        for key, n in zip(('crit', 'two', 'all'), n_emissions_list):
            softdrop[key].append([getECFs_softdrop(**params, n_emissions=n)])
    np.savez(file_path,
             ungroomed_c1s=np.asarray(ungroomed[0]),
             ungroomed_c1s_two=np.asarray(ungroomed[1]),
             ungroomed_c1s_all=np.asarray(ungroomed[2]),
             rss_c1s_crit=np.asarray(rss['crit']),
             rss_c1s_critsub=np.asarray(rss['critsub']),
             rss_c1s_precrit=np.asarray(rss['precrit']),
             rss_c1s_precritsub=np.asarray(rss['precritsub']),
             rss_c1s_two=np.asarray(rss['two']),
             rss_c1s_all=np.asarray(rss['all']),
             softdrop_c1s_crit=np.asarray(softdrop['crit']),
             softdrop_c1s_two=np.asarray(softdrop['two']),
             softdrop_c1s_all=np.asarray(softdrop['all']))
