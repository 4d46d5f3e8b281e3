"""GPU parity tests: the CUDA path (through the C ABI of libjmc_b200.so and the
Python mirror of the reference API) against the oracle and the golden vectors
generated from the unmodified reference.

Bars (BASELINE.json north_star):
  * fixed-sample replay: weights <= 1e-6 relative (we hold 1e-10), NaN pattern
    identical, bin assignment bit-exact;
  * counts are integers and compared exactly; FP64 histogram sums differ from
    NumPy's only by summation order (NumPy: diff of a cumsum over sorted
    blocks; here: atomics), tolerance stated per test.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu

SEED = 20260101
JETS = (('q', 'quark'), ('g', 'gluon'))
W_RTOL = 1e-10          # the contract is 1e-6; FP64 in NumPy order does far better


def close(actual, desired, rtol=W_RTOL, atol=0.):
    """allclose with identical NaN / inf pattern"""
    actual, desired = np.asarray(actual, dtype=float), np.asarray(desired, dtype=float)
    assert_array_equal(np.isnan(actual), np.isnan(desired))
    assert_allclose(actual, desired, rtol=rtol, atol=atol, equal_nan=True)


def ulp_close(actual, desired, ulps):
    actual, desired = np.asarray(actual, dtype=float), np.asarray(desired, dtype=float)
    assert_array_equal(np.isnan(actual), np.isnan(desired))
    ok = np.abs(actual - desired) <= ulps * np.spacing(np.abs(desired))
    assert (ok | np.isnan(desired)).all(), "max ulp diff %g" % np.nanmax(
        np.abs(actual - desired) / np.spacing(np.abs(desired)))


# ---------------------------------------------------------------------------
def test_philox_matches_oracle(cuda):
    import gpu_helpers as H
    for seed, first, n, k in ((0, 0, 1000, 2), (SEED, 2 ** 33 + 5, 777, 6), (2 ** 63 + 11, 2 ** 40, 64, 5)):
        assert_array_equal(H.uniforms(n, seed, first, k), O.philox_uniforms53(seed, first, n, k))


def test_kats(cuda, golden):
    """SURVEY 8(c) known answers through the Python mirror"""
    from jetmontecarlo_b200.analytics import qcd_utils as Q
    from jetmontecarlo_b200.analytics.radiators import fixedcoupling as FC
    from jetmontecarlo_b200.analytics.radiators import running_coupling as RC
    from jetmontecarlo_b200.numerics import observables as OBS
    from jetmontecarlo_b200.numerics import weights as W
    g = golden("kats")
    kat = dict(zip(g['names'].tolist(), g['values'].tolist()))
    assert Q.alpha_fixed == kat['alpha_fixed'] and Q.beta_0 == kat['beta_0'] and Q.MU_NP == kat['MU_NP']
    mine = {'alpha_s(.2,.3)': Q.alpha_s(.2, .3), 'alpha_s(1e-3,1e-3)': Q.alpha_s(1e-3, 1e-3),
            'C_groomed(.2,.3,.1,2)': OBS.C_groomed(.2, .3, .1, 2),
            'C_groomed(.2,.3,.1,2,zpre=.03,f=1,MLL)': OBS.C_groomed(.2, .3, .1, 2, z_pre=.03, f=1, acc='MLL'),
            'C_groomed(.2,.3,.1,2,zpre=.1,f=0,MLL)': OBS.C_groomed(.2, .3, .1, 2, z_pre=.1, f=0, acc='MLL'),
            'C_ungroomed(.2,.3,2,MLL)': OBS.C_ungroomed(.2, .3, 2, 'MLL')}
    c, s = np.array([[.2, .3]]), np.array([[.05, .4]])
    for j, jet in JETS:
        mine.update({
            j + ':splittingFn(.2,LL)': Q.splittingFn(.2, jet, 'LL'),
            j + ':splittingFn(.2,LL_nonsing)': Q.splittingFn(.2, jet, 'LL_nonsing'),
            j + ':splittingFn(.2,MLL)': Q.splittingFn(.2, jet, 'MLL'),
            j + ':radiatorWeight(.2,.3,fc,LL)': W.radiatorWeight(.2, .3, jet, True, 'LL'),
            j + ':radiatorWeight(.2,.3,rc,MLL)': W.radiatorWeight(.2, .3, jet, False, 'MLL'),
            j + ':critPDFAnalytic_fc_LL(.2,.3,.1)': FC.critPDFAnalytic_fc_LL(.2, .3, z_c=.1, jet_type=jet),
            j + ':critRadAnalytic(.3,.1)': RC.critRadAnalytic(.3, .1, jet),
            j + ':critRadAnalytic(2e-3,.1)': RC.critRadAnalytic(2e-3, .1, jet),
            j + ':critRadAnalytic(1e-4,.1)': RC.critRadAnalytic(1e-4, .1, jet),
            j + ':critPDFAnalytic(.2,.3,.1)': RC.critPDFAnalytic(.2, .3, .1, jet),
            j + ':critRadPrimeAnalytic(.3,.1)': RC.critRadPrimeAnalytic(.3, .1, jet),
            j + ':subRadAnalytic_fc_LL(.01,2)': FC.subRadAnalytic_fc_LL(.01, 2, jet),
            j + ':subPDFAnalytic_fc_LL(.01,2)': FC.subPDFAnalytic_fc_LL(.01, 2, jet),
            j + ':subPDFAnalytic_fc_LL(.01,2,R=.3)': FC.subPDFAnalytic_fc_LL(.01, 2, jet, maxRadius=.3),
            j + ':subRadAnalytic(.01,2)': RC.subRadAnalytic(.01, 2, jet),
            j + ':subRadAnalytic(1e-5,2)': RC.subRadAnalytic(1e-5, 2, jet),
            j + ':subRadPrimeAnalytic(.01,2)': RC.subRadPrimeAnalytic(.01, 2, jet),
            j + ':subPDFAnalytic(.01,2,R=.3)': RC.subPDFAnalytic(.01, 2, jet, maxRadius=.3),
            j + ':precriticalEmissionWeight(.03,.3,.1)': W.precriticalEmissionWeight(.03, .3, .1, jet),
            j + ':twoEmissionWeight(fc)': W.twoEmissionWeight(c, s, .1, 2., jet, True)[0],
            j + ':twoEmissionWeight(rc)': W.twoEmissionWeight(c, s, .1, 2., jet, False)[0],
        })
    for name, val in mine.items():
        close(val, kat[name])
    assert np.isnan(mine['q:critRadAnalytic(1e-4,.1)']) and np.isnan(mine['g:subRadAnalytic(1e-5,2)'])


def test_function_vectors(cuda, golden):
    """every analytic function of the path on the reference's own outputs,
    Landau (NaN) regions included"""
    from jetmontecarlo_b200.analytics import qcd_utils as Q
    from jetmontecarlo_b200.analytics.radiators import fixedcoupling as FC
    from jetmontecarlo_b200.analytics.radiators import running_coupling as RC
    from jetmontecarlo_b200.numerics import observables as OBS
    from jetmontecarlo_b200.numerics import weights as W
    g = golden("functions")
    z, th, C, zpre, maxr = g['z'], g['theta'], g['C'], g['z_pre'], g['max_radius']
    close(Q.alpha_s(z, th), g['alpha_s'])
    for j, jet in JETS:
        for acc in ('LL', 'LL_nonsing', 'MLL'):
            close(Q.splittingFn(z, jet, acc), g[f'{j}:splittingFn:{acc}'])
            for fc in (True, False):
                close(W.radiatorWeight(z, th, jet, fc, acc), g[f'{j}:radiatorWeight:{acc}:{int(fc)}'])
        for zc in (.05, .1, .2):
            close(FC.critPDFAnalytic_fc_LL(z, th, zc, jet), g[f'{j}:critPDF_fc:{zc}'])
            close(RC.critPDFAnalytic(z, th, zc, jet), g[f'{j}:critPDF_rc:{zc}'])
            # radiators: cancellation between O(1) terms -> absolute floor
            close(RC.critRadAnalytic(th, zc, jet), g[f'{j}:critRad_rc:{zc}'], atol=1e-12)
            close(RC.critRadPrimeAnalytic(th, zc, jet), g[f'{j}:critRadPrime_rc:{zc}'])
            close(FC.critRadAnalytic_fc_LL(th, zc, jet), g[f'{j}:critRad_fc:{zc}'])
            close(W.precriticalEmissionWeight(zpre, th, zc, jet), g[f'{j}:preWeight:{zc}'])
            close(FC.preRadAnalytic_fc_LL(zpre, th, zc, jet), g[f'{j}:preRad_fc:{zc}'])
        for beta in (2., 3., 1.5, 4.):
            close(FC.subRadAnalytic_fc_LL(C, beta, jet), g[f'{j}:subRad_fc:{beta}'])
            close(FC.subRadPrimeAnalytic_fc_LL(C, beta, jet), g[f'{j}:subRadPrime_fc:{beta}'])
            close(FC.subPDFAnalytic_fc_LL(C, beta, jet), g[f'{j}:subPDF_fc:{beta}'])
            close(FC.subPDFAnalytic_fc_LL(C, beta, jet, maxRadius=maxr), g[f'{j}:subPDF_fc_R:{beta}'])
            close(RC.subRadAnalytic(C, beta, jet), g[f'{j}:subRad_rc:{beta}'], atol=1e-12)
            close(RC.subRadPrimeAnalytic(C, beta, jet), g[f'{j}:subRadPrime_rc:{beta}'])
            close(RC.subPDFAnalytic(C, beta, jet), g[f'{j}:subPDF_rc:{beta}'], rtol=1e-9)
            close(RC.subPDFAnalytic(C, beta, jet, maxRadius=maxr), g[f'{j}:subPDF_rc_R:{beta}'], rtol=1e-9)
            close(RC.subRadAnalytic(C, beta, jet, maxRadius=maxr), g[f'{j}:subRad_rc_R:{beta}'], atol=1e-12)
    zz, zp = np.maximum(z, .1), np.minimum(zpre, .1)
    for beta in (2., 3., 1.5):
        for acc in ('LL', 'MLL'):
            # beta = 2 is a square in NumPy and here: bit-exact.  Other exponents go
            # through pow(), where libm and CUDA may differ in the last place.
            cmp = assert_array_equal if beta == 2. else (lambda a, b: ulp_close(a, b, 4))
            cmp(OBS.C_ungroomed(z, th, beta, acc), g[f'C_ungroomed:{beta}:{acc}'])
            cmp(OBS.C_groomed(zz, th, .1, beta, acc=acc), g[f'C_groomed:{beta}:{acc}'])
            cmp(OBS.C_groomed(zz, th, .1, beta, z_pre=zp, f=1., acc=acc), g[f'C_groomed_pre_f1:{beta}:{acc}'])
            cmp(OBS.C_groomed(zz, th, .1, beta, z_pre=zp, f=.3, acc=acc), g[f'C_groomed_pre_f.3:{beta}:{acc}'])
            cmp(OBS.C_groomed(zz, th, .1, beta, z_pre=.1, f=0., acc=acc), g[f'C_groomed_mmdt:{beta}:{acc}'])


# ---------------------------------------------------------------------------
def test_samplers_match_oracle(cuda):
    """jmc_sample == the reference's lin/log inverse transforms applied to the
    same Philox uniforms; areas and ranges as the reference's own sampler tests
    assert (simple_tests/test_jetSamplers.py:50-171)."""
    import gpu_helpers as H
    from jetmontecarlo_b200 import _lib
    n, seed = 20000, 99
    u = O.philox_uniforms53(seed, 0, n, 2)
    for method, eps in (('lin', 0.), ('log', 1e-10), ('log', 1e-3)):
        s, j = H.sample(_lib.SAMPLER_CRITICAL, method, n, seed, epsilon=eps, zc=.1)
        so, jo, _ = O.sample_critical(method, u, .1, eps)
        close(s, so, rtol=1e-14)
        close(j, jo, rtol=1e-9, atol=1e-300)     # (z - zc) cancels: compare through the same z below
        if method == 'log':
            assert_array_equal(j, (s[:, 0] - .1) * s[:, 1])
        assert s[:, 0].min() > .1 and s[:, 0].max() < .5 and s[:, 1].max() <= 1.
        s, j = H.sample(_lib.SAMPLER_UNGROOMED, method, n, seed, epsilon=eps, radius=.8)
        so, jo, _ = O.sample_ungroomed(method, u, eps, .8)
        close(s, so, rtol=1e-14)
        close(j, jo, rtol=1e-13)
        s, j = H.sample(_lib.SAMPLER_PRECRITICAL, method, n, seed, epsilon=eps, zc=.2)
        so, jo, _ = O.sample_precritical(method, u[:, 0], .2, eps)
        close(s, so, rtol=1e-14)
        s, j = H.sample(_lib.SAMPLER_SIMPLE, method, n, seed, epsilon=eps, lo=.1, hi=.5)
        so, jo, _ = O.sample_simple(method, u[:, 0], (.1, .5), eps)
        close(s, so, rtol=1e-14)
        close(j, jo, rtol=1e-14)


def test_sampler_api(cuda, tmp_path):
    from jetmontecarlo_b200.montecarlo.sampler import set_seed, simpleSampler
    from jetmontecarlo_b200.numerics.radiators.samplers import (criticalSampler, precriticalSampler,
                                                                ungroomedSampler)
    set_seed(5)
    n = 10000
    s = criticalSampler('lin', zc=.1)
    s.generateSamples(n)
    assert s.getSamples().shape == (n, 2) and len(s.jacobians) == n and s.area == (.5 - .1) * 1.
    assert abs(s.samples[:, 0].min() - .1) < .02 * .4 and abs(s.samples[:, 0].max() - .5) < .02 * .4
    with pytest.raises(AttributeError):
        s.generateSamples(3)              # like the reference: samples is an ndarray now
    s.clearSamples()
    s.generateSamples(5)
    assert len(s.samples) == 5 and len(s.jacobians) == n + 5      # jacobians are never cleared
    u = ungroomedSampler('log', epsilon=1e-5)
    u.generateSamples(n)
    assert u.area == np.log(1e5) ** 2.
    assert_array_equal(np.array(u.jacobians), u.samples[:, 0] * u.samples[:, 1])
    u.saveSamples(str(tmp_path / "s.npz"))
    v = ungroomedSampler('log', epsilon=1e-5)
    v.loadSamples(str(tmp_path / "s.npz"))
    assert_array_equal(v.samples, u.samples) and v.area == u.area
    p = precriticalSampler('log', zc=.1, epsilon=1e-8)
    p.generateSamples(n)
    assert p.samples.shape == (n,) and p.samples.max() <= .1 and p.samples.min() >= 1e-9
    q = simpleSampler('lin', bounds=[2, 5])
    q.generateSamples(n)
    assert q.area == 3 and q.samples.min() >= 2 and q.samples.max() <= 5
    set_seed(5)
    s2 = criticalSampler('lin', zc=.1)
    s2.generateSamples(n)
    assert not np.array_equal(s2.samples[:5], s.samples)     # new stream ...
    set_seed(5)
    s3 = criticalSampler('lin', zc=.1)
    s3.generateSamples(n)
    assert_array_equal(s2.samples, s3.samples)               # ... reproducible from the seed


# ---------------------------------------------------------------------------
def _check_integration(g, tag, obs, wj, area, edges, first=None, last=None):
    """hist + finalize through the C ABI and through the integrator mirror"""
    import gpu_helpers as H
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    ref = O.integrate_hist(obs, wj, edges, area, first, last)
    assert_array_equal(H.bin_index(obs, edges), O.bin_index(obs, edges))
    sw, sw2, cnt = H.hist(obs, wj, edges)
    assert_array_equal(cnt, ref['count'])
    # NumPy's sums are differences of a running cumsum: absolute error ~ eps * total
    tot, tot2 = np.nansum(np.abs(wj)), np.nansum(wj ** 2)
    close(sw, ref['sum_w'], rtol=1e-12, atol=1e-13 * tot)
    close(sw2, ref['sum_w2'], rtol=1e-12, atol=1e-13 * tot2)
    it = integrator()
    if first is not None:
        it.setFirstBinBndCondition(first)
    if last is not None:
        it.setLastBinBndCondition(last)
    it.bins, it.hasBins = edges, True
    it.setDensity(obs, wj, area)
    it.integrate()
    widths = np.diff(edges)
    scale = area / len(obs)
    close(it.density, g[tag + 'density'], rtol=1e-12, atol=1e-13 * tot * scale / widths.min())
    close(it.densityErr, g[tag + 'density_err'], rtol=1e-7, atol=1e-7 * np.sqrt(tot2) * scale / widths.min())
    close(np.asarray(it.integral), g[tag + 'integral'], rtol=1e-11, atol=1e-12 * tot * scale)
    close(it.integralErr, g[tag + 'integral_err'], rtol=1e-7, atol=1e-7 * np.sqrt(tot2) * scale)
    assert isinstance(it.integral, list) and len(it.integral) == len(edges) and len(it.integralErr) == len(edges) - 1


@pytest.mark.parametrize("name,method", [("c1_log0", 'log'), ("c1_lin1", 'lin'), ("c1_log2", 'log')])
def test_c1_replay(cuda, golden, name, method):
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    g = golden(name)
    samples, n = g['samples'], len(g['samples'])
    for jet in ('quark', 'gluon'):
        for fc, acc in ((True, 'LL'), (False, 'MLL')):
            tag = f"{jet}_{'fc' if fc else 'rc'}_{acc}:"
            ig = Integrand.radiator(method, float(g['epsilon']), beta=2., jet_type=jet, fixedcoupling=fc,
                                    obs_acc=acc, split_acc=acc)
            edges = g[tag + 'edges']
            jac, obs, wj, bins = H.eval_integrand(ig, n, sub=samples, edges=edges)
            assert_array_equal(jac, g['jac'])
            assert_array_equal(obs, g[tag + 'obs'])                       # bit-exact observable
            close(wj, g[tag + 'weights'] * g['jac'])
            assert_array_equal(bins, O.bin_index(g[tag + 'obs'], edges))  # bit-exact bins
            assert ig.area == float(g['area'])
            # data-dependent edges through the mirror's setBins
            it = integrator()
            it.setBins(100, obs, method)
            assert_array_equal(it.bins, edges)
            assert_array_equal(it.bin_midpoints, g[tag + 'midpoints'])
            _check_integration(g, tag, obs, wj, ig.area, edges, last=[0., 'plus'])


@pytest.mark.parametrize("method,zc", [('log', .1), ('log', .05), ('lin', .1), ('lin', .05)])
def test_critical_replay(cuda, golden, method, zc):
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    g = golden(f"crit_{method}_{zc}")
    samples, n = g['samples'], len(g['samples'])
    for jet in ('quark', 'gluon'):
        for fc in (True, False):
            for beta in (2., 3.):
                tag = f"{jet}_{'fc' if fc else 'rc'}_b{beta:g}:"
                ig = Integrand.critical(method, zc, float(g['epsilon']), beta=beta, jet_type=jet, fixedcoupling=fc)
                edges = g[tag + 'edges']
                jac, obs, wj, bins = H.eval_integrand(ig, n, crit=samples, edges=edges)
                assert_array_equal(jac, g['jac'])
                if beta == 2.:
                    assert_array_equal(obs, g[tag + 'obs'])
                else:
                    ulp_close(obs, g[tag + 'obs'], 4)
                wj_ref = g[tag + 'weights'] * g['jac']
                wj_ref = wj_ref if fc else np.nan_to_num(wj_ref)
                close(wj, wj_ref)
                # bins from OUR observable against the oracle's binning of the reference's observable:
                # identical unless pow() moved a sample across an edge (only possible at the two samples
                # that define the first/last edge, which sit on the edge by construction)
                ref_bins = O.bin_index(g[tag + 'obs'], edges)
                if beta == 2.:
                    assert_array_equal(bins, ref_bins)
                else:
                    assert (bins != ref_bins).sum() <= 2
                _check_integration(g, tag, g[tag + 'obs'], wj_ref, ig.area, edges, last=[1., 'minus'])


def test_c3_replay(cuda, golden):
    """Soft Drop (mMDT form) critical x subsequent, verbatim twoEmissionWeight: the headline integrand"""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    g = golden("c3")
    crit, sub, n = g['crit'], g['sub'], len(g['crit'])
    edges = g['edges_fixed']
    jc, js = g['jac_crit'], g['jac_sub']
    for jet in ('quark', 'gluon'):
        for fc in (True, False):
            for acc in ('LL', 'MLL'):
                tag = f"{jet}_{'fc' if fc else 'rc'}_{acc}:"
                ig = Integrand.crit_sub('log', .1, 1e-10, beta=2., jet_type=jet, fixedcoupling=fc, obs_acc=acc)
                jac, obs, wj, bins = H.eval_integrand(ig, n, crit=crit, sub=sub, edges=edges)
                assert_array_equal(jac, jc * js)
                assert_array_equal(obs, g[tag + 'obs'])
                wj_ref = np.nan_to_num(g[tag + 'weights']) * jc * js
                close(wj, wj_ref, rtol=1e-9)
                assert_array_equal(bins, O.bin_index(g[tag + 'obs'], edges))
                assert ig.area == float(g['area_crit']) * float(g['area_sub'])
                _check_integration(g, tag, obs, wj_ref, ig.area, edges, last=[1., 'minus'])
            # raw weights (NaN kept) through the array API
            ig = Integrand.crit_sub('log', .1, 1e-10, beta=2., jet_type=jet, fixedcoupling=fc, nan_to_num=False)
            _, _, wj, _ = H.eval_integrand(ig, n, crit=crit, sub=sub)
            close(wj, g[f"{jet}_{'fc' if fc else 'rc'}_MLL:weights"] * jc * js, rtol=1e-9)


@pytest.mark.parametrize("jet", ['quark', 'gluon'])
def test_c4_replay(cuda, golden, jet):
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    g3, g = golden("c3"), golden(f"c4_{jet}_fc")
    n = len(g3['crit'])
    ig = Integrand.pre_crit_sub('log', .1, 1e-10, beta=2., jet_type=jet)
    edges = g['edges']
    jac, obs, wj, bins = H.eval_integrand(ig, n, crit=g3['crit'], sub=g3['sub'], pre=g3['z_pre'],
                                          zb_u=g3['zero_bin_u'], edges=edges)
    assert_array_equal(obs, g['obs'])
    assert_array_equal(jac, g['jacs'])
    close(wj, g['weights'] * g['jacs'])
    assert_array_equal(bins, O.bin_index(g['obs'], edges))
    assert ig.area == float(g['area'])
    _check_integration(g, '', obs, g['weights'] * g['jacs'], ig.area, edges, last=[1., 'minus'])


@pytest.mark.parametrize("name,jet,fc,acc,beta", [("c5_quark_fc_b2", 'quark', True, 'LL', 2.),
                                                  ("c5_gluon_rc_b2", 'gluon', False, 'MLL', 2.),
                                                  ("c5_quark_fc_b3", 'quark', True, 'LL', 3.),
                                                  ("c5_gluon_rc_b3", 'gluon', False, 'MLL', 3.)])
def test_c5_grid_replay(cuda, golden, name, jet, fc, acc, beta):
    """gen_crit_sub_num_rad loop body on a theta_crit grid (generation.py:303-335)"""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    g = golden(name)
    samples, n = g['samples'], len(g['samples'])
    for i, thc in enumerate(g['grid']):
        ig = Integrand.radiator('log', 1e-10, beta=beta, jet_type=jet, fixedcoupling=fc, obs_acc=acc,
                                split_acc=acc, theta_scale=thc)
        jac, obs, wj, _ = H.eval_integrand(ig, n, sub=samples)
        obs_ref = O.ecf_ungroomed(samples[:, 0], samples[:, 1] * thc, beta, acc)
        if beta == 2.:
            assert_array_equal(obs, obs_ref)
        else:
            ulp_close(obs, obs_ref, 4)
        assert_array_equal(jac, g['jac'] * thc)
        it = integrator()
        it.setLastBinBndCondition([0., 'plus'])
        it.setBins(40, obs_ref, 'log', min_log_bin=thc ** beta * 1e-15)
        assert_array_equal(it.bins, g['edges'][i])
        it.setDensity(obs_ref, wj, float(g['area']))
        it.integrate()
        close(it.density, g['density'][i], rtol=1e-9, atol=1e-12 * np.abs(g['density'][i]).max())
        close(np.asarray(it.integral), g['integral'][i], rtol=1e-9, atol=1e-12)
        close(it.integralErr, g['integral_err'][i], rtol=1e-7)


@pytest.mark.parametrize("method", ['lin', 'log'])
def test_simple_replay(cuda, golden, method):
    g = golden(f"simple_{method}")
    x = g['samples']
    _check_integration(g, '', x, g['weights'] * g['jac'], float(g['area']), g['edges'], first=0.)


def test_integrate_1d(cuda):
    """test_simpleIntegrator.py known answers: int (n+1) x^n = x^(n+1)"""
    from jetmontecarlo_b200.montecarlo.integrator import integrate_1d
    from jetmontecarlo_b200.montecarlo.sampler import set_seed
    set_seed(3)
    val, err, x = integrate_1d(lambda x: 2. * x, [.1, .5], bin_space='lin', num_samples=200000)
    assert abs(val - (x ** 2 - .1 ** 2)) < 5 * err and err < 2e-3
    integ, err, xs = integrate_1d(lambda x: 3. * x ** 2., [0., 1.], bin_space='lin', num_samples=200000,
                                  num_bins=30, bnd_cond=0., bnd_cond_bin='first')
    assert np.all(np.abs(np.asarray(integ)[1:] - xs[1:] ** 3 + xs[0] ** 3) < 5 * err + 1e-12)


# ---------------------------------------------------------------------------
def test_histogram_semantics(cuda):
    """np.histogram edge cases: last bin closed, out of range and NaN dropped but
    counted in the normalisation, NaN weights poison their bin and all higher
    ones, empty input, ragged sizes, many bins"""
    import gpu_helpers as H
    rng = np.random.RandomState(3)
    edges = np.logspace(-6, 0, 51)
    obs = np.exp(rng.uniform(np.log(1e-7), np.log(2.), 100003))
    obs[:6] = [edges[0], edges[-1], edges[10], np.nan, 0., -1.]
    w = rng.normal(size=obs.size)
    assert_array_equal(H.bin_index(obs, edges), O.bin_index(obs, edges))
    ref = O.histograms(obs, w, edges)
    sw, sw2, cnt = H.hist(obs, w, edges)
    assert_array_equal(cnt, ref[2])
    close(sw, ref[0], rtol=1e-10, atol=1e-10)
    close(sw2, ref[1], rtol=1e-10, atol=1e-10)
    # NaN weights
    w2 = w.copy()
    w2[np.nanargmin(np.abs(obs - 1e-3))] = np.nan
    ref = O.histograms(obs, w2, edges)
    sw, sw2, cnt = H.hist(obs, w2, edges)
    assert_array_equal(np.isnan(sw), np.isnan(ref[0]))
    assert_array_equal(np.isnan(sw2), np.isnan(ref[1]))
    assert np.isnan(ref[0]).any() and not np.isnan(ref[0]).all()
    close(sw, ref[0], rtol=1e-10, atol=1e-10)
    assert_array_equal(cnt, ref[2])
    # NaN weight on a sample below the first edge poisons everything; above the last edge nothing
    w3 = w.copy()
    w3[4] = np.nan
    assert np.isnan(O.histograms(obs, w3, edges)[0]).all() and np.isnan(H.hist(obs, w3, edges)[0]).all()
    w4 = w.copy()
    w4[np.nanargmax(obs)] = np.nan
    assert np.nanmax(obs) > edges[-1]
    assert not np.isnan(O.histograms(obs, w4, edges)[0]).any() and not np.isnan(H.hist(obs, w4, edges)[0]).any()
    # sizes: empty, 1, ragged; 2 edges; 5000 edges (params.py NUM_RAD_BINS)
    for n in (0, 1, 255, 257):
        sw, sw2, cnt = H.hist(obs[6:6 + n], w[6:6 + n], edges)
        assert_array_equal(cnt, O.histograms(obs[6:6 + n], w[6:6 + n], edges)[2])
    for ne in (2, 5000):
        e = np.logspace(-7, .4, ne)
        ref = O.histograms(obs, w, e)
        sw, sw2, cnt = H.hist(obs, w, e)
        assert_array_equal(cnt, ref[2])
        close(sw, ref[0], rtol=1e-10, atol=1e-10)
    # unweighted
    _, _, cnt = H.hist(obs, None, edges)
    assert_array_equal(cnt, O.histograms(obs, w, edges)[2])


def test_integrator_api_errors(cuda):
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    it = integrator()
    with pytest.raises(AssertionError, match="Need bins"):
        it.setDensity(np.ones(3), np.ones(3), 1.)
    it.setBins(5, np.array([.1, .2, .9]), 'lin')
    assert_array_equal(it.bins, np.linspace(0, .9, 5))
    with pytest.raises(AssertionError, match="Need density"):
        it.integrate()
    it.setDensity(np.array([.1, .2, .9]), np.ones(3), 1.)
    with pytest.raises(AssertionError, match="unambiguous"):
        it.integrate()
    it.setFirstBinBndCondition(0.)
    it.setLastBinBndCondition([1., 'minus'])
    with pytest.raises(AssertionError, match="unambiguous"):
        it.integrate()
    it.firstBinBndCond = None
    it.setMonotone()
    it.integrate()
    assert it.hasMCIntegral and set(it.montecarlo_data_dict()) == {
        'bins', 'bin midpoints', 'density', 'density error', 'integral', 'integral error'}
    # callers set density by attribute and integrate (sudakov_utils.py:425-430)
    it2 = integrator()
    it2.bins, it2.hasBins = np.linspace(0, 1, 11), True
    it2.density, it2.densityErr, it2.hasMCDensity = np.full(10, 2.), np.full(10, .1), True
    it2.setFirstBinBndCondition(0.)
    it2.integrate()
    assert_allclose(it2.integral, np.linspace(0, 2, 11), rtol=1e-14)
    assert_allclose(it2.integralErr, np.linspace(.01, .1, 10), rtol=1e-14)
    # setBins on the whole (N,2) sample array runs min/max over both columns (generation.py:103)
    it3 = integrator()
    s = np.array([[.2, .7], [.3, .01]])
    it3.setBins(4, s, 'log')
    assert_array_equal(it3.bins, np.logspace(np.log10(.01), np.log10(.7), 4))


# ---------------------------------------------------------------------------
def _oracle_fused(kind, method, eps, jet, fc, acc, n, seed, first=0, zc=.1, beta=2.):
    """the oracle's view of what the fused FP64 kernel does: Philox uniforms ->
    reference sampler formulas -> reference observable / weight"""
    k = {'radiator': 2, 'crit': 2, 'crit_sub': 4, 'pre_crit_sub': 6}[kind]
    u = O.philox_uniforms53(seed, first, n, k)
    if kind == 'radiator':
        s, jac, area = O.sample_ungroomed(method, u, eps)
        obs = O.ecf_ungroomed(s[:, 0], s[:, 1], beta, acc)
        w = O.weight_radiator(s[:, 0], s[:, 1], jet, fc, acc)
        return obs, w * jac, area
    c, jc, ac = O.sample_critical(method, u[:, :2], zc, eps)
    if kind == 'crit':
        obs = O.ecf_groomed(c[:, 0], c[:, 1], zc, beta)
        w = O.weight_critical(c[:, 0], c[:, 1], zc, jet, fc)
        return obs, (w if fc else np.nan_to_num(w)) * jc, ac
    s, js, as_ = O.sample_ungroomed(method, u[:, 2:4], eps)
    if kind == 'crit_sub':
        csub = s[:, 0] * s[:, 1] ** beta * s[:, 1] ** beta
        obs = np.maximum(O.ecf_groomed(c[:, 0], c[:, 1], zc, beta, z_pre=zc, f=0., acc=acc), csub)
        w = np.nan_to_num(O.weight_two_emission(c, s, zc, beta, jet, fc))
        return obs, w * jc * js, ac * as_
    p, jp, ap = O.sample_precritical(method, u[:, 4], zc, eps)
    th_c = c[:, 1]
    cdf_eps = np.exp(-2. * O.casimir(jet) * O.ALPHA_FIXED / np.pi * np.log(O.R0 / th_c) * np.log(zc / eps))
    zp = p.copy()
    zp[u[:, 5] < cdf_eps] = 0.
    c_sub = s[:, 0] * th_c ** beta
    obs = np.maximum(O.ecf_groomed(c[:, 0], th_c, zc, beta, z_pre=zp, f=1., acc='LL'), c_sub)
    w = (O.weight_critical(c[:, 0], th_c, zc, jet, True) * O.pdf_sub_fc(c_sub / th_c ** beta, beta, jet)
         * O.weight_precritical_zerobin(zp, th_c, zc, jet, eps))
    return obs, w * np.where(zp == 0, jc * js, jc * jp * js), ac * ap * as_


@pytest.mark.parametrize("kind,method,jet,fc,acc", [
    ('radiator', 'log', 'quark', True, 'LL'),       # C1
    ('radiator', 'log', 'gluon', False, 'MLL'),     # C2 (integration form)
    ('radiator', 'lin', 'quark', False, 'MLL'),
    ('crit', 'log', 'quark', True, 'LL'),
    ('crit', 'log', 'gluon', False, 'LL'),
    ('crit_sub', 'log', 'quark', False, 'MLL'),     # C3 headline (verbatim rc)
    ('crit_sub', 'log', 'quark', True, 'MLL'),      # C3 fc control
    ('crit_sub', 'lin', 'gluon', True, 'LL'),
    ('pre_crit_sub', 'log', 'quark', True, 'LL'),   # C4
    ('pre_crit_sub', 'log', 'gluon', True, 'LL'),
])
def test_fused_f64_matches_oracle(cuda, kind, method, jet, fc, acc):
    """jmc_run in FP64 mode == oracle applied to the same Philox stream: counts
    exactly, weighted sums to summation-order rounding; first_index offsets and
    accumulation over several launches give the same job."""
    import gpu_helpers as H
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.fused import Integrand
    eps = 1e-10 if method == 'log' else 0.
    n, seed = 60000, 1234567
    obs, wj, area = _oracle_fused(kind, method, eps, jet, fc, acc, n, seed)
    make = {'radiator': lambda: Integrand.radiator(method, eps, beta=2., jet_type=jet, fixedcoupling=fc, obs_acc=acc,
                                                   split_acc=acc),
            'crit': lambda: Integrand.critical(method, .1, eps, jet_type=jet, fixedcoupling=fc),
            'crit_sub': lambda: Integrand.crit_sub(method, .1, eps, jet_type=jet, fixedcoupling=fc, obs_acc=acc),
            'pre_crit_sub': lambda: Integrand.pre_crit_sub(method, .1, eps, jet_type=jet)}[kind]
    ig = make()
    assert ig.area == area
    # edges that keep the extreme samples strictly inside the outer bins: the device and NumPy exp/log may
    # differ in the last place, which must not decide whether the min/max sample is dropped
    lo = max(np.nanmin(obs) * .999, 1e-15)
    edges = (np.logspace(np.log10(lo), np.log10(np.nanmax(obs) * 1.001), 100) if method == 'log'
             else np.linspace(0, np.nanmax(obs) * 1.001, 100))
    ref = O.integrate_hist(obs, wj, edges, area, last_bc=[1., 'minus'])
    r = H.run(ig, n, seed, edges, 'f64', want_minmax=True)
    assert_array_equal(r['count'].cpu().numpy(), ref['count'])
    tot, tot2 = np.abs(wj).sum(), (wj ** 2).sum()
    close(r['sum_w'].cpu().numpy(), ref['sum_w'], rtol=1e-9, atol=1e-12 * tot)
    close(r['sum_w2'].cpu().numpy(), ref['sum_w2'], rtol=1e-9, atol=1e-12 * tot2)
    close(r['minmax'], [np.min(obs), np.max(obs)], rtol=1e-14)
    d, de, integ, ie = H.finalize(r['sum_w'], r['sum_w2'], r['count'], edges, area, n, _lib.BC_LAST_MINUS, 1.)
    scale = area / n / np.diff(edges).min()
    close(d, ref['density'], rtol=1e-9, atol=1e-12 * tot * scale)
    close(integ, ref['integral'], rtol=1e-9, atol=1e-11 * tot * area / n)
    # split into two launches with index offsets, accumulating
    acc_t = None
    for first, cnt in ((0, 25000), (25000, 35000)):
        rr = H.run(ig, cnt, seed, edges, 'f64', first=first,
                   accumulate_into=None if acc_t is None else acc_t)
        acc_t = (rr['sum_w'], rr['sum_w2'], rr['count'])
    assert_array_equal(acc_t[2].cpu().numpy(), ref['count'])
    close(acc_t[0].cpu().numpy(), ref['sum_w'], rtol=1e-9, atol=1e-12 * tot)


def test_fused_nan_poison(cuda):
    """nan_to_num off: running-coupling NaN weights poison the histogram like NumPy"""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    n, seed = 20000, 5
    u = O.philox_uniforms53(seed, 0, n, 2)
    c, jc, area = O.sample_critical('log', u, .1, 1e-10)
    obs = O.ecf_groomed(c[:, 0], c[:, 1], .1, 2.)
    wj = O.weight_critical(c[:, 0], c[:, 1], .1, 'quark', False) * jc
    edges = np.logspace(-22, np.log10(.5), 60)
    ref = O.histograms(obs, wj, edges)
    ig = Integrand.critical('log', .1, 1e-10, fixedcoupling=False, nan_to_num=False)
    r = H.run(ig, n, seed, edges, 'f64')
    assert_array_equal(np.isnan(r['sum_w'].cpu().numpy()), np.isnan(ref[0]))
    assert np.isnan(ref[0]).any()
    assert_array_equal(r['count'].cpu().numpy(), ref[2])


def test_fused_integrator_api(cuda):
    """setBinsFused / setDensityFused / fused.integrate == array API on the same samples"""
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    n, seed = 50000, 77
    obs, wj, area = _oracle_fused('crit', 'log', 1e-10, 'quark', True, 'LL', n, seed)
    ig = Integrand.critical('log', .1, 1e-10, jet_type='quark', fixedcoupling=True)
    it = fused.integrate(ig, n, numbins=100, binspacing='log', seed=seed, precision='f64', last_bc=[1., 'minus'])
    edges, _ = O.make_edges(100, obs, 'log')
    close(it.bins, edges, rtol=1e-13)
    ref = O.integrate_hist(obs, wj, it.bins, area, last_bc=[1., 'minus'])
    close(it.density, ref['density'], rtol=1e-9, atol=1e-9 * np.abs(ref['density']).max())
    close(np.asarray(it.integral), ref['integral'], rtol=1e-9, atol=1e-12)
    # critical Sudakov against the closed form of the reference's test (critSudakov_fc_LL KAT region):
    # the integral is a cdf between 0 and 1 and monotone
    integ = np.asarray(it.integral)
    assert integ[-1] == 1. and np.all(np.diff(integ) >= -1e-12) and integ[0] > 0
