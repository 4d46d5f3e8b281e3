"""Drop-in on the GPU: scripts written against the REFERENCE's import paths (``from jetmontecarlo.... import *``) run on
the device path after ``jetmontecarlo_b200.install_as('jetmontecarlo')``.  The flows below use the reference's API in
the order its own tests do -- examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py:124-220 (one critical
emission, log sampling, Sudakov factor against the closed form) and
examples/tests/multipleEmissions_tests/LL/test_critsub_sudakov.py:151-258 (critical + subsequent) -- minus the
plotting, and check the Monte Carlo cdfs against the closed forms those tests plot them against.  A third flow runs
the production chain of examples/sudakov_comparisons: numerical radiator on a (z_pre, theta_crit) grid -> table ->
per-event conditional inverse-transform samples (sudakov_utils.py:134-182, 568-607)."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_PRELUDE = """
import numpy as np
from jetmontecarlo.montecarlo.integrator import *
from jetmontecarlo.numerics.radiators.samplers import *
from jetmontecarlo.numerics.observables import *
from jetmontecarlo.numerics.radiators.generation import *
from jetmontecarlo.analytics.sudakov_factors.fixedcoupling import *
"""

_ONE_EMISSION = _PRELUDE + """
results = {}
testInt = integrator()
testInt.setLastBinBndCondition([1., 'minus'])
for jet_type in ['quark', 'gluon']:
    for beta in betas:
        for zc in zcuts:
            for eps in epsilons:
                testSampler = criticalSampler('log', zc=zc, epsilon=eps)
                testSampler.generateSamples(numSamples)
                samples = testSampler.getSamples()
                z = samples[:, 0]; theta = samples[:, 1]
                obs = C_groomed(z, theta, zc, beta)
                testInt.setBins(numBins, obs, 'log')
                weights = criticalEmissionWeight(z, theta, zc, jet_type, fixedcoupling=True)
                jacs = testSampler.jacobians
                area = testSampler.area
                testInt.setDensity(obs, weights * jacs, area)
                testInt.integrate()
                xs = testInt.bins[:-1]
                analytic = np.array([float(a) for a in critSudakov_fc_LL(xs, zc, beta, jet_type=jet_type)])
                results[(jet_type, beta, zc, eps)] = (np.array(testInt.integral[:-1]), np.array(testInt.integralErr), analytic)
"""

_CRIT_SUB = _PRELUDE + """
results = {}
test_int = integrator()
test_int.setLastBinBndCondition([1., 'minus'])
for jet_type in ['quark', 'gluon']:
    for beta in betas:
        for z_c in z_cuts:
            for eps in epsilons:
                crit_sampler = criticalSampler('log', zc=z_c, epsilon=eps)
                crit_sampler.generateSamples(NUM_SAMPLES)
                samples = crit_sampler.getSamples()
                z_crit = samples[:, 0]
                theta_crit = samples[:, 1]
                sub_sampler = ungroomedSampler('log', epsilon=eps)
                sub_sampler.generateSamples(NUM_SAMPLES)
                samples = sub_sampler.getSamples()
                c_sub = samples[:, 0]
                obs = (C_groomed(z_crit, theta_crit, z_c, beta, z_pre=0., f=1., acc='LL') + c_sub)
                test_int.setBins(NUM_BINS, obs, 'log')
                weights = (criticalEmissionWeight(z_crit, theta_crit, z_c, jet_type, fixedcoupling=True)
                           * subPDFAnalytic_fc_LL(c_sub, beta, jet_type=jet_type))
                jacs = (np.array(crit_sampler.jacobians) * np.array(sub_sampler.jacobians))
                area = (np.array(crit_sampler.area) * np.array(sub_sampler.area))
                test_int.setDensity(obs, weights * jacs, area)
                test_int.integrate()
                results[(jet_type, beta, z_c, eps)] = (np.array(test_int.bins), np.array(test_int.integral),
                                                       np.array(test_int.integralErr))
"""


@pytest.fixture()
def as_jetmontecarlo():
    import jetmontecarlo_b200
    saved = {k: v for k, v in sys.modules.items() if k == "jetmontecarlo" or k.startswith("jetmontecarlo.")}
    for k in saved:
        del sys.modules[k]
    jetmontecarlo_b200.install_as("jetmontecarlo")
    yield
    for k in [k for k in sys.modules if k == "jetmontecarlo" or k.startswith("jetmontecarlo.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_one_emission_flow_of_the_reference_tests(cuda, as_jetmontecarlo):
    from jetmontecarlo_b200.montecarlo.sampler import set_seed
    set_seed(20260101)
    ns = dict(numSamples=400_000, numBins=60, betas=[2., 3.], zcuts=[.05, .1], epsilons=[1e-5, 1e-10])
    exec(_ONE_EMISSION, ns)
    assert ns['integrator'].__module__ == 'jetmontecarlo_b200.montecarlo.integrator'
    assert len(ns['results']) == 16
    for key, (integral, err, analytic) in ns['results'].items():
        # the Sudakov factor of one critical emission against its closed form (what the reference plots): where the
        # Monte Carlo has support (the closed form's power-law tail below the sampled range is cut by epsilon)
        ok = (analytic > .05) & (err > 0)
        pull = (integral[ok] - analytic[ok]) / err[ok]
        # the phase space below epsilon is an offset of the cdf (up to ~0.1 at 1e-5): invisible at 1e-10 only
        worst = np.abs(integral[ok] - analytic[ok]).max()
        assert ok.sum() > 15 and worst < (.15 if key[3] > 1e-8 else 3e-2), (key, worst)
        # (away from the end point, where the boundary condition pins the integral and the error bar collapses)
        body = analytic[ok] < .97
        assert key[3] > 1e-8 or np.abs(pull[body]).max() < 5., (key, np.abs(pull[body]).max())


def test_crit_sub_flow_of_the_reference_tests(cuda, as_jetmontecarlo):
    """critical + subsequent emission with fixed coupling: the cdf of C_crit + C_sub ends at 1, is monotone, and lies
    below the single-emission Sudakov factor (a second emission can only raise the observable)"""
    from jetmontecarlo_b200.montecarlo.sampler import set_seed
    from jetmontecarlo_b200.analytics.sudakov_factors.fixedcoupling import critSudakov_fc_LL
    set_seed(7)
    ns = dict(NUM_SAMPLES=400_000, NUM_BINS=60, betas=[2.], z_cuts=[.1], epsilons=[1e-5, 1e-10])
    exec(_CRIT_SUB, ns)
    for (jet_type, beta, z_c, eps), (bins, integral, err) in ns['results'].items():
        assert integral[-1] == 1. and np.all(np.diff(integral) > -4 * np.r_[err, err[-1]][:-1] - 1e-9)
        single = np.array([float(a) for a in critSudakov_fc_LL(bins[:-1], z_c, beta, jet_type=jet_type)])
        ok = (single > .05) & (single < .999)
        assert np.all(integral[:-1][ok] <= single[ok] + 4 * err[ok] + 2e-3), (jet_type, eps)
        assert (integral[:-1][ok] < single[ok]).mean() > .8


def test_production_chain_radiator_table_to_conditional_samples(cuda, as_jetmontecarlo, tmp_path):
    """gen_pre_num_rad on the device (2-D integrator) -> stored in the file catalog as the reference's
    phase_space_sampling.py stores it -> loaded as a TableRadiator -> one z_pre per critical angle by the conditional
    inverse-transform kernel: the samples follow the cdf the table defines"""
    from jetmontecarlo.numerics.radiators.samplers import criticalSampler, precriticalSampler
    from jetmontecarlo.numerics.radiators.generation import gen_pre_num_rad
    from jetmontecarlo_b200.montecarlo.sampler import set_seed
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    set_seed(3)
    z_cut = .1
    crit, pre = criticalSampler('log', zc=z_cut, epsilon=1e-8), precriticalSampler('log', zc=z_cut, epsilon=1e-8)
    crit.generateSamples(2_000_000)
    pre.generateSamples(2_000_000)
    fn, d = gen_pre_num_rad(pre, crit, jet_type='quark', splitfn_accuracy='LL', bin_space='log', fixed_coupling=True,
                            num_bins=40)
    from file_management.data_management import save_new_data
    from file_management import catalog_utils
    from jetmontecarlo_b200.file_management.data_management import load_and_tabulate
    catalog_utils.set_output_root(tmp_path / "output")
    params = {'z_cut': z_cut, 'jet type': 'quark', 'fixed coupling': True, 'number of samples': 2_000_000}
    save_new_data({'bins': np.array(d['bins']), 'integral': np.array(d['integral'])}, 'numerical integral',
                  'pre-critical radiator', params, '.npz')
    rad = load_and_tabulate('pre-critical radiator', params)
    assert isinstance(rad, S.TableRadiator) and rad.zs.shape == (40, 40)
    theta = np.full(40_000, .3)
    z_pre, w = S.generate_precritical_samples(rad, theta, z_cut, seed=5)
    z_pre = z_pre.cpu().numpy()
    assert np.all(w.cpu().numpy() == 1.) and np.all((z_pre >= -1e-6) & (z_pre <= z_cut * 1.001))
    # (uniforms below the cdf's value at the first probe point are extrapolated linearly, as interp1d does in the
    # reference: slightly negative z_pre for the no-emission events)
    # empirical cdf of the samples against exp(-R(z, theta)) of the same table
    zs = np.sort(z_pre[z_pre > 1e-7])
    probe = zs[::400]
    emp = np.searchsorted(np.sort(z_pre), probe, side='right') / len(z_pre)
    model = np.exp(-rad(probe, np.full_like(probe, .3)))
    assert np.abs(emp - model).max() < .03, np.abs(emp - model).max()


def test_production_route_phase_space_to_final_pdf(cuda, as_jetmontecarlo, tmp_path):
    """The reference's production route for the groomed distribution, end to end on GPU outputs and through its file
    catalog: phase-space sampling -> numerical radiators + normalised splitting function, stored
    (examples/event_generation/phase_space_sampling.py:95-300) -> load_radiators / load_splittingfns
    (file_management/load_data.py:114-195) -> inverse-transform Sudakov samples per event, stored and loaded
    (examples/sudakov_comparisons/sudakov_utils.py:100-238) -> get_mc_all (:568-607, groomer 'mmdt').
    Every stage's samples are compared with the closed form the reference plots that stage against (fixed coupling,
    LL: exp(-R) of critRadAnalytic_fc_LL, subRadAnalytic_fc_LL, preRadAnalytic_fc_LL)."""
    from jetmontecarlo.numerics.radiators.samplers import criticalSampler, precriticalSampler, ungroomedSampler
    from jetmontecarlo.numerics.radiators.generation import gen_numerical_radiator, gen_pre_num_rad, gen_crit_sub_num_rad
    from jetmontecarlo.numerics.splitting import gen_normalized_splitting
    from jetmontecarlo.numerics.observables import C_groomed
    from jetmontecarlo.montecarlo.sampler import simpleSampler
    from jetmontecarlo.utils.hist_utils import vals_to_pdf
    from file_management.data_management import save_new_data, load_data
    from file_management.load_data import load_radiators, load_splittingfns, load_sudakov_samples
    from file_management import catalog_utils
    from jetmontecarlo_b200.montecarlo.sampler import set_seed
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    catalog_utils.set_output_root(tmp_path / "output")
    set_seed(11)
    z_cut, beta, eps, jet, f_soft = .1, 2., 1e-10, 'quark', 1.
    n_phase, n_events, rad_bins = 2_000_000, 400_000, 80
    common = {'jet type': jet, 'fixed coupling': True, 'observable accuracy': 'LL', 'splitting function accuracy': 'LL',
              'epsilon': eps, 'bin space': 'log', 'number of MC events': n_phase}
    rad_params = dict(common, **{'number of radiator bins': rad_bins})
    split_params = dict(common, **{'number of splitting function bins': 40})

    # ---- phase_space_sampling.py ----
    split_fn = gen_normalized_splitting(n_phase, z_cut, jet_type=jet, accuracy='LL', fixed_coupling=True, num_bins=40)
    save_new_data(split_fn, 'serialized function', 'splitting function', dict(split_params, z_cut=z_cut), '.pkl')
    crit_sampler = criticalSampler('log', zc=z_cut, epsilon=eps)
    crit_sampler.generateSamples(n_phase)
    save_new_data(crit_sampler, 'montecarlo samples', 'critical phase space', dict(common, z_cut=z_cut), '.pkl')
    _, crit_rad = gen_numerical_radiator(crit_sampler, 'crit', jet, obs_accuracy='LL', splitfn_accuracy='LL', beta=None,
                                         bin_space='log', fixed_coupling=True, num_bins=rad_bins)
    save_new_data(crit_rad, 'numerical integral', 'critical radiator', dict(rad_params, z_cut=z_cut), '.npz')
    pre_sampler = precriticalSampler('log', zc=z_cut, epsilon=eps)
    pre_sampler.generateSamples(n_phase)
    _, pre_rad = gen_pre_num_rad(pre_sampler, crit_sampler, jet, splitfn_accuracy='LL', bin_space='log',
                                 fixed_coupling=True, num_bins=rad_bins)
    save_new_data(pre_rad, 'numerical integral', 'pre-critical radiator', dict(rad_params, z_cut=z_cut), '.npz')
    sub_sampler = ungroomedSampler('log', epsilon=eps)
    sub_sampler.generateSamples(n_phase)
    _, sub_rad = gen_crit_sub_num_rad(sub_sampler, jet, obs_accuracy='LL', splitfn_accuracy='LL', beta=beta,
                                      epsilon=eps, bin_space='log', fixed_coupling=True, num_bins=rad_bins)
    save_new_data(sub_rad, 'numerical integral', 'subsequent radiator', dict(rad_params, beta=beta), '.npz')

    # ---- what a later job finds in the catalog ----
    stored = load_data('montecarlo samples', 'critical phase space', dict(common, z_cut=z_cut))
    np.testing.assert_array_equal(stored.getSamples(), crit_sampler.getSamples())
    radiators = load_radiators(rad_params, [z_cut], [beta], as_tables=True)
    splitting_functions = load_splittingfns(split_params, [z_cut])
    assert isinstance(radiators['subsequent'][beta], S.TableRadiator) and callable(radiators['critical'][z_cut])

    # ---- sudakov_utils.generate_*_samples: one sample per event from cdfs exp(-R) ----
    theta_crits, theta_w = S.generate_critical_samples(radiators['critical'][z_cut], n_events, seed=21)
    # (a critical angle beyond the last node of a table meets the interpolant's linear extrapolation -- a slightly
    # negative radiator, a cdf above 1 that falls -- and raises here as it does in the reference; the handful of such
    # events, theta_crit within 1e-4 of 1, condition on the last node instead)
    pre_table, sub_table = radiators['pre-critical'][z_cut], radiators['subsequent'][beta]
    c_subs, c_w = S.generate_subsequent_samples(sub_table, np.clip(theta_crits, sub_table.ys[0], sub_table.ys[-1]), beta,
                                                seed=22)
    z_pres, z_w = S.generate_precritical_samples(pre_table, np.clip(theta_crits, pre_table.ys[0], pre_table.ys[-1]),
                                                 z_cut, seed=23)
    sud_params = dict(rad_params, **{'number of events': n_events})
    save_new_data({'samples': theta_crits, 'weights': theta_w}, 'montecarlo samples',
                  'critical sudakov inverse transform', dict(sud_params, z_cut=z_cut), '.npz')
    save_new_data({'samples': c_subs, 'weights': c_w}, 'montecarlo samples',
                  'subsequent sudakov inverse transform', dict(sud_params, z_cut=z_cut, beta=beta), '.npz')
    save_new_data({'samples': z_pres, 'weights': z_w}, 'montecarlo samples',
                  'pre-critical sudakov inverse transform', dict(sud_params, z_cut=z_cut), '.npz')
    inv = load_sudakov_samples(sud_params, [z_cut], [beta])

    # ---- get_mc_all(z_cut, beta, groomer='mmdt') ----
    theta_crits = np.asarray(inv['critical'][z_cut][beta]['samples'])
    c_subs = np.asarray(inv['subsequent'][z_cut][beta]['samples'])
    w_all = (np.asarray(inv['critical'][z_cut][beta]['weights']) * np.asarray(inv['subsequent'][z_cut][beta]['weights']))
    assert np.asarray(inv['pre-critical'][z_cut]['samples']).shape == theta_crits.shape
    z_sampler = simpleSampler('lin', bounds=[z_cut, .5])          # getLinSample(z_cut, 1/2) per event
    z_sampler.generateSamples(n_events)
    z_crits = np.asarray(z_sampler.getSamples()).reshape(-1)
    assert z_crits.min() >= z_cut and z_crits.max() <= .5
    c_crits = np.nan_to_num(np.asarray(C_groomed(z_crits, theta_crits, z_cut, beta, z_pre=z_cut, f=f_soft, acc='LL')))
    obs = np.maximum(c_crits, c_subs)
    weights = np.asarray(splitting_functions[z_cut](z_crits, theta_crits)) * w_all
    xs, pdf = vals_to_pdf(vals=obs, num_bins=60, weights=weights, log_cutoff=eps)
    assert np.all(np.isfinite(pdf)) and np.all(pdf >= 0) and pdf.sum() > 0
    # vals_to_pdf returns d sigma / d log10 C on its bins (zero bin first): back to the mass of every bin
    bins = np.insert(np.logspace(np.log10(eps), 0, 60), 0, 1e-100)
    cdf = np.cumsum(pdf / (xs * np.log(10)) * np.diff(bins))
    assert abs(cdf[-1] - 1.) < 1e-9

    # ---- every stage of the route against the closed forms the reference plots it against ----
    from jetmontecarlo.analytics.radiators.fixedcoupling import (critRadAnalytic_fc_LL, subRadAnalytic_fc_LL,
                                                                 preRadAnalytic_fc_LL)
    # critical angles: P(theta_crit <= t) = exp(-R_crit(t)), the events without a critical emission sit at 0
    t_grid = np.logspace(-8, -.05, 25)
    emp = np.searchsorted(np.sort(theta_crits), t_grid, side='right') / len(theta_crits)
    model = np.exp(-np.asarray(critRadAnalytic_fc_LL(t_grid, z_cut, jet), dtype=float))
    assert np.abs(emp - model).max() < .02, np.abs(emp - model).max()
    # subsequent emissions of the events in a narrow band of critical angles: P(C_sub <= c | theta) = exp(-R_sub(c; theta))
    band = (theta_crits > .3) & (theta_crits < .33)
    assert band.sum() > 2000
    c_grid = np.logspace(-7, np.log10(.3 ** beta / 2.), 20)
    emp = np.searchsorted(np.sort(c_subs[band]), c_grid, side='right') / band.sum()
    model = np.exp(-np.asarray(subRadAnalytic_fc_LL(c_grid, beta, jet, maxRadius=.315), dtype=float))
    assert np.abs(emp - model).max() < .05, np.abs(emp - model).max()
    # pre-critical energy fractions in the same band: P(z_pre <= z | theta) = exp(-R_pre(z, theta))
    z_pres = np.asarray(inv['pre-critical'][z_cut]['samples'])
    z_grid = np.logspace(-7, np.log10(z_cut * .98), 15)
    emp = np.searchsorted(np.sort(z_pres[band]), z_grid, side='right') / band.sum()
    model = np.exp(-np.asarray(preRadAnalytic_fc_LL(z_grid, .315, z_cut, jet), dtype=float))
    assert np.abs(emp - model).max() < .05, np.abs(emp - model).max()
    # the final distribution: a cdf that rises over the sampled range
    assert np.all(np.diff(cdf) >= -1e-12) and cdf[1] < .5 and cdf[-2] > .95
