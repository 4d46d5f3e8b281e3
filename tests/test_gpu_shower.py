"""GPU tests of the veto-algorithm shower kernel (BASELINE configs 2 and 3-iii).

Parity chain:
  reference (gen_jets + getECFs_*, MT19937)  == oracle.shower_emissions on the same MT stream   [CPU, golden]
  oracle.shower_emissions on a Philox stream == FP64 kernel, jet by jet                          [here]
  FP32 kernel                                 ~ FP64 kernel, statistically                        [here]
plus the closed-form LL angularity cdf the reference's own shower test plots against
(examples/tests/partonshower_tests/test_partonshower_angularities.py:87-98).
"""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_jets(n, seed, jet, acc, beta=2., first=0):
    """chains from the oracle fed with each jet's Philox stream"""
    chains, trials = [], []
    for i in range(n):
        u = O.philox_uniforms53(seed, first + i, 1, 1200)[0]
        it = iter(u)
        count = [0]

        def nxt():
            count[0] += 1
            return next(it)
        chains.append(O.shower_emissions(nxt, beta, jet, acc, azimuth_draw=False))
        trials.append(count[0])
    return chains, trials


@pytest.mark.parametrize("jet,acc", [('quark', 'LL'), ('quark', 'MLL'), ('gluon', 'LL'), ('gluon', 'MLL')])
def test_shower_f64_jet_by_jet(cuda, jet, acc):
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    n, seed = 300, 424242
    chains, n_uniform = _oracle_jets(n, seed, jet, acc)
    jets = PS.gen_jets(n, beta=2., jet_type=jet, acc=acc, verbose=0, seed=seed)
    n_em, n_tr = jets.emission_counts()
    assert_array_equal(n_em, [len(c) for c in chains])
    per_trial = 2 if acc == 'LL' else 3
    assert_array_equal(n_tr * per_trial, n_uniform)
    for oacc in ('LL', 'MLL'):
        for nem in (1, 2, 'all'):
            ref = [O.ecf_from_chain_ungroomed(c, 2., oacc, nem) for c in chains]
            got = PS.getECFs_ungroomed(jets, beta=2., obs_acc=oacc, n_emissions=nem)
            assert_allclose(got, ref, rtol=1e-12)
            ref = [O.ecf_from_chain_softdrop(c, .1, 2., 0., oacc, nem) for c in chains]
            got = PS.getECFs_softdrop(jets, .1, beta=2., beta_sd=0., obs_acc=oacc, n_emissions=nem)
            assert_allclose(got, ref, rtol=1e-12)
        ref = [O.ecf_from_chain_softdrop(c, .1, 2., 1., oacc, 1) for c in chains]
        assert_allclose(PS.getECFs_softdrop(jets, .1, beta=2., beta_sd=1., obs_acc=oacc, n_emissions=1), ref, rtol=1e-12)
    # the event record itself
    lens, zs, ths = jets.chains(max_emissions=32)
    assert_array_equal(lens, n_em)
    for i in (0, 1, 17, n - 1):
        ref_chain = np.array(chains[i]).reshape(-1, 2)
        assert_allclose(zs[i, :lens[i]], ref_chain[:, 0], rtol=1e-13)
        assert_allclose(ths[i, :lens[i]], ref_chain[:, 1], rtol=1e-13)
        assert np.isnan(zs[i, lens[i]:]).all()
    # mean multiplicities of the survey (section 6): 2.6 / 3.9 emissions per quark jet (LL / MLL), ~30 MLL trials
    if jet == 'quark':
        assert abs(n_em.mean() - (2.6 if acc == 'LL' else 3.9)) < .5


def test_shower_histogram_and_fused_api(cuda):
    """fused histogram of the shower == np.histogram of the per-jet observables; index offsets; weights = counts"""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    n, seed = 20000, 7
    edges = np.insert(np.logspace(-10, 0, 100), 0, 1e-100)      # plot_comparisons.py:296-298 (zero bin first)
    ig = Integrand.shower(beta=2., jet_type='quark', acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL',
                          n_emissions=2)
    jets = PS.gen_jets(n, beta=2., jet_type='quark', acc='MLL', verbose=0, seed=seed)
    obs = np.array(PS.getECFs_softdrop(jets, .1, beta=2., obs_acc='MLL', n_emissions=2))
    r = H.run(ig, n, seed, edges, 'f64', want_minmax=True)
    ref = np.histogram(obs, edges)[0]
    assert_array_equal(r['count'].cpu().numpy(), ref)
    assert_array_equal(r['sum_w'].cpu().numpy(), ref.astype(float))
    assert_array_equal(r['sum_w2'].cpu().numpy(), ref.astype(float))
    assert_allclose(r['minmax'], [obs.min(), obs.max()], rtol=1e-14)
    acc = None
    for first, cnt in ((0, 5000), (5000, 15000)):
        rr = H.run(ig, cnt, seed, edges, 'f64', first=first, accumulate_into=acc)
        acc = (rr['sum_w'], rr['sum_w2'], rr['count'])
    assert_array_equal(acc[2].cpu().numpy(), ref)


def test_shower_ll_closed_form_and_f32(cuda):
    """LL shower with fixed coupling: cdf of the leading-emission angularity is the LL Sudakov
    exp(-C_R alpha/(beta pi) ln^2(2 C)) (test_partonshower_angularities.py:87-98); FP32 and FP64 kernels agree
    with it and with each other within binomial errors"""
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    n = 2_000_000
    edges = np.logspace(-8, np.log10(.5), 60)
    for jet in ('quark', 'gluon'):
        ig = Integrand.shower(beta=2., jet_type=jet, acc='LL', obs_acc='LL', n_emissions=1)
        cdfs = {}
        for prec, seed in (('f64', 3), ('f32', 4)):
            it = fused.integrate(ig, n, edges=edges, seed=seed, precision=prec, first_bc=0.)
            # jets are unweighted: density * width * N = counts
            counts = np.asarray(it.density) * np.diff(edges) * n
            below = n - counts.sum() - 0      # jets under the first edge (above the last there are none: C <= 1/2)
            cdfs[prec] = (below + np.cumsum(counts)) / n
        cr = O.casimir(jet)
        exact = np.exp(-cr * O.ALPHA_FIXED / (2. * np.pi) * np.log(2. * edges[1:]) ** 2.)
        sigma = np.sqrt(exact * (1 - exact) / n) + 1e-9
        for prec in cdfs:
            assert np.abs((cdfs[prec] - exact) / sigma).max() < 5., prec
        assert np.abs(cdfs['f32'] - cdfs['f64']).max() < 5 * np.sqrt(2) * sigma.max()


def test_shower_mll_f32_vs_f64(cuda):
    """MLL veto shower, Soft Drop two-emission ECF (BASELINE config 3-iii): FP32 vs FP64 pdfs within 3 sigma"""
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    n = 3_000_000
    edges = np.insert(np.logspace(-10, np.log10(.5), 80), 0, 1e-100)
    for jet in ('quark', 'gluon'):
        ig = Integrand.shower(beta=2., jet_type=jet, acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL',
                              n_emissions=2)
        a = fused.integrate(ig, n, edges=edges, seed=11, precision='f32', first_bc=0.)
        b = fused.integrate(ig, n, edges=edges, seed=12, precision='f64', first_bc=0.)
        err = np.sqrt(a.densityErr ** 2 + b.densityErr ** 2)
        m = err > 0
        pulls = (a.density[m] - b.density[m]) / err[m]
        assert np.abs(pulls).max() < 5. and (np.abs(pulls) > 3).sum() <= 2
        assert abs(np.asarray(a.integral)[-1] - 1.) < 1e-9 and abs(np.asarray(b.integral)[-1] - 1.) < 1e-9


@pytest.mark.parametrize("jet", ['quark', 'gluon'])
def test_c2_veto_shower_against_integrated_radiator(cuda, jet):
    """BASELINE config 2 at full size (1e8 jets): the two forms of the ungroomed MLL calculation must agree.
    Veto form: cdf of the leading-emission angularity of the MLL shower.  Integration form: exp(-R(C)) with
    R(C) the Monte Carlo integral of radiatorWeight (running coupling, reduced splitting function) over
    z theta^beta > C.  (The reference validates its shower the same way:
    examples/tests/partonshower_tests/test_partonshower_angularities.py:180-209.)"""
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    n_jets, n_samples = 100_000_000, 2_000_000_000
    edges = np.logspace(-7, np.log10(.5), 50)
    sh = Integrand.shower(beta=2., jet_type=jet, acc='MLL', obs_acc='LL', n_emissions=1)
    it = fused.integrate(sh, n_jets, edges=edges, seed=21, precision='f32', first_bc=0.)
    counts = np.asarray(it.density) * np.diff(edges) * n_jets
    cdf = ((n_jets - counts.sum()) + np.cumsum(counts)) / n_jets          # P(C_lead < edge_k), k = 1..
    rad = Integrand.radiator('log', 1e-10, beta=2., jet_type=jet, fixedcoupling=False, obs_acc='LL', split_acc='MLL')
    rt = fused.integrate(rad, n_samples, edges=edges, seed=22, precision='f32', last_bc=[0., 'plus'])
    R, dR = np.asarray(rt.integral)[1:], np.append(rt.integralErr[1:], 0.)
    sud = np.exp(-R)
    sigma = np.sqrt(cdf * (1 - cdf) / n_jets + (sud * dR) ** 2) + 2e-5
    assert np.abs((cdf - sud) / sigma).max() < 5., np.abs((cdf - sud) / sigma).max()
    assert np.abs(cdf - sud).max() < 1e-3
