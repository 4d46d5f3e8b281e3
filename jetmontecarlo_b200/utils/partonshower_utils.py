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
from ..analytics.qcd_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..analytics.qcd_utils import MU_NP, check_jet_type
from ..fused import PRECISIONS, Integrand
from ..montecarlo.sampler import next_key


CHAIN_RECORD = 64        # emissions per jet in a first event record (MLL gluon jets average about 20)


# ---- the event record as objects, on request ------------------------------------------------------------------------
# With split_soft=False (the only mode gen_jets uses, :321) a jet IS its ordered (z, theta) chain: every splitting
# takes the hard daughter of the previous one.  The device produces chains; the classes below rebuild the reference's
# tree of Parton objects with momenta from a chain, for the scripts that walk it (shower visualisation, custom
# groomers).  ref: utils/partonshower_utils.py:13-112, utils/vector_utils.py:22-137
class Vector:
    """A three-vector with the few operations the shower record needs (magnitude, rotation about an axis)."""

    def __init__(self, vector):
        self.vector = np.array(vector, dtype=float)

    def dim(self):
        return len(self.vector)

    def mag(self):
        return float(np.sqrt(np.dot(self.vector, self.vector)))

    def unit(self):
        return Vector(self.vector / self.mag())

    def rotate_around(self, axis, angle):
        """Rodrigues rotation of this vector about ``axis`` by ``angle``"""
        k = axis.vector / axis.mag()
        v = self.vector
        return Vector(v * np.cos(angle) + np.cross(k, v) * np.sin(angle) + k * np.dot(k, v) * (1. - np.cos(angle)))


def angle(vec1, vec2):
    """opening angle of two vectors"""
    cos = np.dot(vec1.vector, vec2.vector) / (vec1.mag() * vec2.mag())
    return float(np.arccos(np.clip(cos, -1., 1.)))


def rand_perp_vector(vector, rng=None):
    """a unit vector perpendicular to ``vector`` with a uniformly random azimuth"""
    rng = np.random if rng is None else rng
    v = vector.vector / vector.mag()
    helper = np.array([1., 0., 0.]) if abs(v[0]) < .9 else np.array([0., 1., 0.])
    e1 = np.cross(v, helper)
    e1 /= np.sqrt(np.dot(e1, e1))
    e2 = np.cross(v, e1)
    phi = rng.uniform(0., 2. * np.pi)
    return Vector(np.cos(phi) * e1 + np.sin(phi) * e2)


def split_momentum(mom_vector, z, theta, rng=None):
    """(soft, hard) momenta of a splitting with momentum fraction z and opening angle theta"""
    assert mom_vector.dim() == 3, "Currently only splitting 3D vectors!"
    assert mom_vector.mag() > 0, "Cannot split zero momentum."
    perp = rand_perp_vector(mom_vector, rng)
    q = Vector(mom_vector.vector * z).rotate_around(perp, theta / 2)
    k = Vector(mom_vector.vector * (1 - z)).rotate_around(perp, -theta / 2)
    return q, k


class Parton():
    """Momentum, mother / daughters and identifiers of one parton of the shower record."""

    def __init__(self, momentum):
        self.momentum = momentum
        self.mass = None
        self.mother = None
        self.daughter1 = None
        self.daughter2 = None
        self.pid = None
        self.isFinalState = True
        self.type = None

    def split(self, z, theta, rng=None):
        q, k = split_momentum(self.momentum, z, theta, rng)
        self.daughter1, self.daughter2 = Parton(q), Parton(k)
        self.daughter1.mother = self.daughter2.mother = self
        self.isFinalState = False


class Jet:
    """A list of partons built up by a shower."""

    def __init__(self, momentum, radius, partons=None):
        assert momentum.mag() > 0, "Unphysical transverse momentum."
        assert radius > 0, "Unphysical jet radius."
        self.momentum = momentum
        self.R = radius
        self.partons = [Parton(momentum)] if partons is None else list(partons)
        self.has_showered = False

    @classmethod
    def from_chain(cls, chain, radius=1., rng=None):
        """The record the reference's recursive shower leaves behind for the emission chain [(z, theta), ...]:
        mother, then for every emission its soft and hard daughters, the hard one splitting next."""
        from ..analytics.qcd_utils import P_T
        jet = cls(Vector([0, P_T, 0]), radius)
        parton = jet.partons[0]
        parton.type = 'mother'
        for z, theta in chain:
            parton.split(float(z), float(theta), rng)
            parton.daughter1.type, parton.daughter2.type = 'soft', 'hard'
            jet.partons.append(parton.daughter1)
            jet.partons.append(parton.daughter2)
            parton = parton.daughter2
        jet.has_showered = True
        return jet

    def chain(self):
        """the (z, theta) chain back from the tree: z = |soft| / |mother|, theta = angle between the daughters"""
        out, parton = [], self.partons[0]
        while parton.daughter1 is not None:
            out.append((parton.daughter1.momentum.mag() / parton.momentum.mag(),
                        angle(parton.daughter1.momentum, parton.daughter2.momentum)))
            parton = parton.daughter2
        return out


def angularity_shower(parton, ang_init, z_init, theta_init, beta, jet_type, jet, acc='LL', split_soft=True,
                      cutoff=MU_NP, seed=None):
    """Showers ``parton`` inside ``jet`` (both are modified in place, as in the reference): the chain is generated on
    the device starting from angularity ``ang_init`` and attached to the record.  Only the hard daughter splits
    further (``split_soft=False``, what gen_jets uses); a shower of both daughters is a tree the device kernel does
    not generate.  ref: utils/partonshower_utils.py:224-285"""
    if split_soft:
        raise NotImplementedError("angularity_shower: split_soft=True is not on the accelerated path "
                                  "(gen_jets uses split_soft=False)")
    if not z_init * theta_init > cutoff:
        return
    radius = (2. * ang_init) ** (1. / beta)            # the chain of a jet of this radius starts at ang_init
    lens, z, theta = JetList(1, beta, jet_type, radius, acc, cutoff,
                             next_key() if seed is None else int(seed)).chains(CHAIN_RECORD)
    for k in range(int(lens[0])):
        parton.split(float(z[0, k]), float(theta[0, k]))
        parton.daughter1.type, parton.daughter2.type = 'soft', 'hard'
        jet.partons.append(parton.daughter1)
        jet.partons.append(parton.daughter2)
        parton = parton.daughter2


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
        chains = D.full((n, int(max_emissions), 2), float('nan'))
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

    def jets(self, max_emissions=CHAIN_RECORD, rng=None):
        """The jets as ``Jet`` objects (Parton trees with momenta), rebuilt on the host from the device's event record:
        for scripts that walk the tree.  The observables never need this."""
        lens, z, theta = self.chains(max_emissions)
        assert len(lens) == 0 or int(lens.max()) <= max_emissions, "emission chains were truncated: raise max_emissions"
        return [Jet.from_chain(zip(z[i, :lens[i]], theta[i, :lens[i]]), self.radius, rng) for i in range(len(lens))]

    def __iter__(self):
        return iter(self.jets())

    def __getitem__(self, i):
        return self.jets()[i]

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
    out = D.empty((lens.numel(),))
    s = ig.c_struct()
    rc = _lib.lib.jmc_chain_ecf(C.byref(s), D.ptr(lens), D.ptr(chains), lens.numel(), int(chains.shape[1]),
                                D.ptr(out), D.stream())
    # (the kernel checks every record's length against its slots: no separate reduction over ``lens``)
    assert rc != -4, "emission chains were truncated: raise max_emissions"
    _lib.check(rc)
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
