"""The oracle (oracle/jmc_oracle.py) against vectors produced by executing the
unmodified reference (oracle/make_golden.py).  Bit-for-bit: the oracle uses the
same NumPy ufuncs in the same order, so ``assert_array_equal`` (NaN == NaN)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle import jmc_oracle as O

JETS = (('q', 'quark'), ('g', 'gluon'))
SEED = 20260101


def test_kats(golden):
    k = golden("kats")
    kat = dict(zip(k['names'].tolist(), k['values'].tolist()))
    assert kat['alpha_fixed'] == O.ALPHA_FIXED == 0.0785101418528313
    assert kat['beta_0'] == O.BETA_0 == 0.6100939485189322
    assert kat['MU_NP'] == O.MU_NP
    assert kat['alpha_s(.2,.3)'] == O.alpha_running(.2, .3) == 0.10747697766038962
    assert kat['alpha_s(1e-3,1e-3)'] == O.alpha_running(1e-3, 1e-3) == 0.33693459053455094
    # SURVEY.md 8(c) literal values
    assert kat['q:splittingFn(.2,LL)'] == 13.333333333333332
    assert kat['q:splittingFn(.2,MLL)'] == 12.666666666666664
    assert kat['q:radiatorWeight(.2,.3,fc,LL)'] == 1.1106913029976886
    assert kat['q:critPDFAnalytic(.2,.3,.1)'] == 1.2291682807149376
    assert kat['q:twoEmissionWeight(fc)'] == 1.0546664905025822
    assert kat['q:twoEmissionWeight(rc)'] == 2.1045424549030836
    assert kat['landau'] == 2.9276463899588305e-05
    c, s = np.array([[.2, .3]]), np.array([[.05, .4]])
    for j, jet in JETS:
        assert kat[j + ':splittingFn(.2,LL)'] == O.splitting(.2, jet, 'LL')
        assert kat[j + ':splittingFn(.2,LL_nonsing)'] == O.splitting(.2, jet, 'LL_nonsing')
        assert kat[j + ':splittingFn(.2,MLL)'] == O.splitting(.2, jet, 'MLL')
        assert kat[j + ':radiatorWeight(.2,.3,fc,LL)'] == O.weight_radiator(.2, .3, jet, True, 'LL')
        assert kat[j + ':radiatorWeight(.2,.3,rc,MLL)'] == O.weight_radiator(.2, .3, jet, False, 'MLL')
        assert kat[j + ':critPDFAnalytic_fc_LL(.2,.3,.1)'] == O.pdf_crit_fc(.2, .3, .1, jet)
        assert kat[j + ':critRadAnalytic(.3,.1)'] == O.rad_crit_rc(.3, .1, jet)
        assert kat[j + ':critRadAnalytic(2e-3,.1)'] == O.rad_crit_rc(2e-3, .1, jet)
        assert np.isnan(kat[j + ':critRadAnalytic(1e-4,.1)']) and np.isnan(O.rad_crit_rc(1e-4, .1, jet))
        assert kat[j + ':critPDFAnalytic(.2,.3,.1)'] == O.pdf_crit_rc(.2, .3, .1, jet)
        assert kat[j + ':critRadPrimeAnalytic(.3,.1)'] == O.radprime_crit_rc(.3, .1, jet)
        assert kat[j + ':subRadAnalytic_fc_LL(.01,2)'] == O.rad_sub_fc(.01, 2, jet)
        assert kat[j + ':subPDFAnalytic_fc_LL(.01,2)'] == O.pdf_sub_fc(.01, 2, jet)
        assert kat[j + ':subPDFAnalytic_fc_LL(.01,2,R=.3)'] == O.pdf_sub_fc(.01, 2, jet, .3)
        assert kat[j + ':subRadAnalytic(.01,2)'] == O.rad_sub_rc(.01, 2, jet)
        assert np.isnan(kat[j + ':subRadAnalytic(1e-5,2)']) and np.isnan(O.rad_sub_rc(1e-5, 2, jet))
        assert kat[j + ':subRadPrimeAnalytic(.01,2)'] == O.radprime_sub_rc(.01, 2, jet)
        assert kat[j + ':subPDFAnalytic(.01,2,R=.3)'] == O.pdf_sub_rc(.01, 2, jet, .3)
        assert kat[j + ':precriticalEmissionWeight(.03,.3,.1)'] == O.weight_precritical(.03, .3, .1, jet)
        assert kat[j + ':twoEmissionWeight(fc)'] == O.weight_two_emission(c, s, .1, 2., jet, True)[0]
        assert kat[j + ':twoEmissionWeight(rc)'] == O.weight_two_emission(c, s, .1, 2., jet, False)[0]
    assert kat['C_groomed(.2,.3,.1,2)'] == O.ecf_groomed(.2, .3, .1, 2) == 0.01111111111111111
    assert kat['C_groomed(.2,.3,.1,2,zpre=.03,f=1,MLL)'] == O.ecf_groomed(.2, .3, .1, 2, .03, 1, 'MLL')
    assert kat['C_groomed(.2,.3,.1,2,zpre=.1,f=0,MLL)'] == O.ecf_groomed(.2, .3, .1, 2, .1, 0, 'MLL')
    assert kat['C_ungroomed(.2,.3,2,MLL)'] == O.ecf_ungroomed(.2, .3, 2, 'MLL')


def test_closed_form_critical_sudakov(golden):
    """critSudakov_fc_LL (mpmath hyp2f1): the closed form the reference plots its critical-emission MC against"""
    g = golden("kats")
    kat = dict(zip(g['names'].tolist(), g['values'].tolist()))
    assert O.sudakov_crit_fc(.01, .1, 2, 'quark')[0] == kat['q:critSudakov_fc_LL(.01,.1,2)']
    assert O.sudakov_crit_fc(.01, .1, 2, 'gluon')[0] == kat['g:critSudakov_fc_LL(.01,.1,2)']
    # and it IS the cdf of (z - z_c) theta^beta under critPDFAnalytic_fc_LL (vectorised MC on the oracle's samplers)
    n = 1000000
    s, jac, area = O.sample_critical('log', O.mt_uniforms(1, 2 * n).reshape(n, 2), .1, 1e-10)
    w = O.pdf_crit_fc(s[:, 0], s[:, 1], .1, 'quark') * jac
    edges = np.logspace(-8, np.log10(.5), 30)
    r = O.integrate_hist((s[:, 0] - .1) * s[:, 1] ** 2., w, edges, area, last_bc=[1., 'minus'])
    pull = (r['integral'][:-1] - O.sudakov_crit_fc(edges[:-1], .1, 2., 'quark')) / r['integral_err']
    assert np.abs(pull).max() < 4.


def test_function_vectors(golden):
    g = golden("functions")
    z, th, C, zpre, maxr = g['z'], g['theta'], g['C'], g['z_pre'], g['max_radius']
    assert_array_equal(g['alpha_s'], O.alpha_running(z, th))
    nan_seen = 0
    for j, jet in JETS:
        for acc in ('LL', 'LL_nonsing', 'MLL'):
            assert_array_equal(g[f'{j}:splittingFn:{acc}'], O.splitting(z, jet, acc))
            for fc in (True, False):
                assert_array_equal(g[f'{j}:radiatorWeight:{acc}:{int(fc)}'],
                                   O.weight_radiator(z, th, jet, fc, acc))
        for zc in (.05, .1, .2):
            assert_array_equal(g[f'{j}:critPDF_fc:{zc}'], O.pdf_crit_fc(z, th, zc, jet))
            assert_array_equal(g[f'{j}:critPDF_rc:{zc}'], O.pdf_crit_rc(z, th, zc, jet))
            assert_array_equal(g[f'{j}:critRad_rc:{zc}'], O.rad_crit_rc(th, zc, jet))
            nan_seen += int(np.isnan(g[f'{j}:critRad_rc:{zc}']).sum())
            assert_array_equal(g[f'{j}:critRadPrime_rc:{zc}'], O.radprime_crit_rc(th, zc, jet))
            assert_array_equal(g[f'{j}:critRad_fc:{zc}'], O.rad_crit_fc(th, zc, jet))
            assert_array_equal(g[f'{j}:preWeight:{zc}'], O.weight_precritical(zpre, th, zc, jet))
            assert_array_equal(g[f'{j}:preRad_fc:{zc}'], O.rad_pre_fc(zpre, th, zc, jet))
        for beta in (2., 3., 1.5, 4.):
            assert_array_equal(g[f'{j}:subRad_fc:{beta}'], O.rad_sub_fc(C, beta, jet))
            assert_array_equal(g[f'{j}:subRadPrime_fc:{beta}'], O.radprime_sub_fc(C, beta, jet))
            assert_array_equal(g[f'{j}:subPDF_fc:{beta}'], O.pdf_sub_fc(C, beta, jet))
            assert_array_equal(g[f'{j}:subPDF_fc_R:{beta}'], O.pdf_sub_fc(C, beta, jet, maxr))
            assert_array_equal(g[f'{j}:subRad_rc:{beta}'], O.rad_sub_rc(C, beta, jet))
            assert_array_equal(g[f'{j}:subRadPrime_rc:{beta}'], O.radprime_sub_rc(C, beta, jet))
            assert_array_equal(g[f'{j}:subPDF_rc:{beta}'], O.pdf_sub_rc(C, beta, jet))
            assert_array_equal(g[f'{j}:subPDF_rc_R:{beta}'], O.pdf_sub_rc(C, beta, jet, maxr))
            assert_array_equal(g[f'{j}:subRad_rc_R:{beta}'], O.rad_sub_rc(C, beta, jet, maxr))
    assert nan_seen > 100  # the Landau region is covered
    zz = np.maximum(z, .1)
    zp = np.minimum(zpre, .1)
    for beta in (2., 3., 1.5):
        for acc in ('LL', 'MLL'):
            assert_array_equal(g[f'C_ungroomed:{beta}:{acc}'], O.ecf_ungroomed(z, th, beta, acc))
            assert_array_equal(g[f'C_groomed:{beta}:{acc}'], O.ecf_groomed(zz, th, .1, beta, acc=acc))
            assert_array_equal(g[f'C_groomed_pre_f1:{beta}:{acc}'],
                               O.ecf_groomed(zz, th, .1, beta, zp, 1., acc))
            assert_array_equal(g[f'C_groomed_pre_f.3:{beta}:{acc}'],
                               O.ecf_groomed(zz, th, .1, beta, zp, .3, acc))
            assert_array_equal(g[f'C_groomed_mmdt:{beta}:{acc}'],
                               O.ecf_groomed(zz, th, .1, beta, .1, 0., acc))


def _check_integrator(g, tag, obs, wj, area, spacing=None, edges=None, first=None, last=None,
                      min_log_bin=O.MIN_LOG_BIN, numbins=100):
    if edges is None:
        edges, _ = O.make_edges(numbins, obs, spacing, min_log_bin)
    assert_array_equal(g[tag + 'edges'], edges)
    r = O.integrate_hist(obs, wj, edges, area, first, last)
    assert_array_equal(g[tag + 'density'], r['density'])
    assert_array_equal(g[tag + 'density_err'], r['density_err'])
    assert_array_equal(g[tag + 'integral'], r['integral'])
    assert_array_equal(g[tag + 'integral_err'], r['integral_err'])
    # per-sample statement of np.histogram's binning
    idx = O.bin_index(obs, edges)
    assert_array_equal(np.bincount(idx[idx >= 0], minlength=len(edges) - 1), r['count'])


@pytest.mark.parametrize("name,method,case", [("c1_log0", 'log', 0), ("c1_lin1", 'lin', 1),
                                              ("c1_log2", 'log', 2)])
def test_c1_replay(golden, name, method, case):
    g = golden(name)
    n = len(g['samples'])
    u = O.mt_uniforms(SEED + case, 2 * n).reshape(n, 2)
    samples, jac, area = O.sample_ungroomed(method, u, float(g['epsilon']))
    assert_array_equal(samples, g['samples'])
    assert_array_equal(jac, g['jac'])
    assert area == float(g['area'])
    for jet in ('quark', 'gluon'):
        for fc, acc in ((True, 'LL'), (False, 'MLL')):
            tag = f"{jet}_{'fc' if fc else 'rc'}_{acc}:"
            obs = O.ecf_ungroomed(samples[:, 0], samples[:, 1], 2., acc)
            w = O.weight_radiator(samples[:, 0], samples[:, 1], jet, fc, acc)
            assert_array_equal(obs, g[tag + 'obs'])
            assert_array_equal(w, g[tag + 'weights'])
            _check_integrator(g, tag, obs, w * jac, area, spacing=method, last=[0., 'plus'])
            _, mids = O.make_edges(100, obs, method)
            assert_array_equal(mids, g[tag + 'midpoints'])


@pytest.mark.parametrize("method,zc,case", [('log', .1, 10), ('log', .05, 11),
                                            ('lin', .1, 12), ('lin', .05, 13)])
def test_critical_replay(golden, method, zc, case):
    g = golden(f"crit_{method}_{zc}")
    n = len(g['samples'])
    u = O.mt_uniforms(SEED + case, 2 * n).reshape(n, 2)
    samples, jac, area = O.sample_critical(method, u, zc, float(g['epsilon']))
    assert_array_equal(samples, g['samples'])
    assert_array_equal(jac, g['jac'])
    assert area == float(g['area'])
    for jet in ('quark', 'gluon'):
        for fc in (True, False):
            for beta in (2., 3.):
                tag = f"{jet}_{'fc' if fc else 'rc'}_b{beta:g}:"
                obs = O.ecf_groomed(samples[:, 0], samples[:, 1], zc, beta)
                w = O.weight_critical(samples[:, 0], samples[:, 1], zc, jet, fc)
                assert_array_equal(obs, g[tag + 'obs'])
                assert_array_equal(w, g[tag + 'weights'])
                wj = w * jac if fc else np.nan_to_num(w * jac)
                _check_integrator(g, tag, obs, wj, area, spacing=method, last=[1., 'minus'])


def _c3_samples(g):
    n = len(g['crit'])
    u = O.mt_uniforms(SEED + 30, 6 * n)
    crit, jc, ac = O.sample_critical('log', u[:2 * n].reshape(n, 2), .1, 1e-10)
    sub, js, as_ = O.sample_ungroomed('log', u[2 * n:4 * n].reshape(n, 2), 1e-10)
    pre, jp, ap = O.sample_precritical('log', u[4 * n:5 * n], .1, 1e-10)
    zb = u[5 * n:]
    return crit, jc, ac, sub, js, as_, pre, jp, ap, zb


def test_c3_replay(golden):
    g = golden("c3")
    crit, jc, ac, sub, js, as_, pre, jp, ap, zb = _c3_samples(g)
    assert_array_equal(crit, g['crit'])
    assert_array_equal(sub, g['sub'])
    assert_array_equal(pre, g['z_pre'])
    assert_array_equal(zb, g['zero_bin_u'])
    assert_array_equal(jc, g['jac_crit'])
    assert_array_equal(js, g['jac_sub'])
    assert_array_equal(jp, g['jac_pre'])
    assert (ac, as_, ap) == (float(g['area_crit']), float(g['area_sub']), float(g['area_pre']))
    edges = np.logspace(np.log10(1e-10) - 1, np.log10(.5), 100)
    csub = sub[:, 0] * sub[:, 1] ** 2. * sub[:, 1] ** 2.
    assert_array_equal(csub, g['csub'])
    for jet in ('quark', 'gluon'):
        for fc in (True, False):
            w = O.weight_two_emission(crit, sub, .1, 2., jet, fc)
            for acc in ('LL', 'MLL'):
                tag = f"{jet}_{'fc' if fc else 'rc'}_{acc}:"
                assert_array_equal(w, g[tag + 'weights'])
                ccrit = O.ecf_groomed(crit[:, 0], crit[:, 1], .1, 2., z_pre=.1, f=0., acc=acc)
                assert_array_equal(ccrit, g[f'c_crit:{acc}'])
                obs = np.maximum(ccrit, csub)
                assert_array_equal(obs, g[tag + 'obs'])
                _check_integrator(g, tag, obs, np.nan_to_num(w) * jc * js, ac * as_,
                                  edges=edges, last=[1., 'minus'])
    # the survey's caveat: almost all running-coupling weights are NaN -> 0
    assert np.isnan(g['quark_rc_MLL:weights']).mean() > .98


@pytest.mark.parametrize("jet", ['quark', 'gluon'])
def test_c4_replay(golden, jet):
    g3, g = golden("c3"), golden(f"c4_{jet}_fc")
    crit, jc, ac, sub, js, as_, pre, jp, ap, zb = _c3_samples(g3)
    th_c = crit[:, 1]
    cr = O.casimir(jet)
    cdf_eps = np.exp(-2. * cr * O.ALPHA_FIXED / np.pi * np.log(O.R0 / th_c) * np.log(.1 / 1e-10))
    zp = pre.copy()
    zp[zb < cdf_eps] = 0.
    assert_array_equal(zp, g['z_pre_zb'])
    c_sub = sub[:, 0] * th_c ** 2.
    obs = np.maximum(O.ecf_groomed(crit[:, 0], th_c, .1, 2., z_pre=zp, f=1., acc='LL'), c_sub)
    assert_array_equal(obs, g['obs'])
    w = (O.weight_critical(crit[:, 0], th_c, .1, jet, True)
         * O.pdf_sub_fc(c_sub / th_c ** 2., 2., jet)
         * O.weight_precritical_zerobin(zp, th_c, .1, jet, 1e-10))
    assert_array_equal(w, g['weights'])
    jacs = np.where(zp == 0, jc * js, jc * jp * js)
    assert_array_equal(jacs, g['jacs'])
    _check_integrator(g, '', obs, w * jacs, ac * ap * as_, spacing='log', last=[1., 'minus'])


@pytest.mark.parametrize("name,jet,fc,oacc,beta", [("c5_quark_fc_b2", 'quark', True, 'LL', 2.),
                                                   ("c5_gluon_rc_b2", 'gluon', False, 'MLL', 2.),
                                                   ("c5_quark_fc_b3", 'quark', True, 'LL', 3.),
                                                   ("c5_gluon_rc_b3", 'gluon', False, 'MLL', 3.)])
def test_c5_grid_replay(golden, name, jet, fc, oacc, beta):
    g = golden(name)
    n = len(g['samples'])
    u = O.mt_uniforms(SEED + 50, 2 * n).reshape(n, 2)
    samples, jac, area = O.sample_ungroomed('log', u, 1e-10)
    assert_array_equal(samples, g['samples'])
    grid = O.lin_log_mixed_list(1e-10, 1., 12)
    assert_array_equal(grid, g['grid'])
    for i, thc in enumerate(grid):
        th_em = samples[:, 1] * thc
        obs = O.ecf_ungroomed(samples[:, 0], th_em, beta, oacc)
        w = O.weight_radiator(samples[:, 0], th_em, jet, fc, oacc)
        edges, _ = O.make_edges(40, obs, 'log', thc ** beta * 1e-15)
        assert_array_equal(edges, g['edges'][i])
        r = O.integrate_hist(obs, w * (jac * thc), edges, area, last_bc=[0., 'plus'])
        assert_array_equal(r['density'], g['density'][i])
        assert_array_equal(r['density_err'], g['density_err'][i])
        assert_array_equal(r['integral'], g['integral'][i])
        assert_array_equal(r['integral_err'], g['integral_err'][i])


@pytest.mark.parametrize("method,case", [('lin', 60), ('log', 61)])
def test_simple_sampler_replay(golden, method, case):
    g = golden(f"simple_{method}")
    n = len(g['samples'])
    x, jac, area = O.sample_simple(method, O.mt_uniforms(SEED + case, n), (0, 1), float(g['epsilon']))
    assert_array_equal(x, g['samples'])
    assert_array_equal(jac, g['jac'])
    assert area == float(g['area'])
    w = 3. * x ** 2.
    _check_integrator(g, '', x, w * jac, area, spacing=method, first=0., numbins=50)


def test_integrate_1d_replay(golden):
    g = golden("integrate_1d_lin")
    x, jac, area = O.sample_simple('lin', O.mt_uniforms(SEED + 62, 2000), [.1, .5], None)
    edges, _ = O.make_edges(2, x, 'lin')
    r = O.integrate_hist(x, 2. * x * jac, edges, area, first_bc=0)
    assert r['integral'][1] == float(g['value'])
    assert r['integral_err'][0] == float(g['error'])
    g = golden("integrate_1d_log")
    x, jac, area = O.sample_simple('log', O.mt_uniforms(SEED + 63, 2000), [.1, .5], 1e-6)
    edges, _ = O.make_edges(20, x, 'log')
    assert_array_equal(edges, g['xs'])
    r = O.integrate_hist(x, (1. / x) * jac, edges, area, first_bc=0)
    assert_array_equal(r['integral'], g['integral'])
    assert_array_equal(r['integral_err'], g['error'])


@pytest.mark.parametrize("jet,acc,case", [('quark', 'LL', 70), ('quark', 'MLL', 71),
                                          ('gluon', 'LL', 72), ('gluon', 'MLL', 73)])
def test_shower_replay(golden, jet, acc, case):
    g = golden(f"shower_{jet}_{acc}")
    import random
    gen = random.Random(SEED + case)
    lens = g['lens']
    zs, ts = [], []
    res = {k: [] for k in g if ':' in k}
    for nj in range(len(lens)):
        chain = O.shower_emissions(gen.random, 2., jet, acc)
        assert len(chain) == lens[nj]
        zs += [c[0] for c in chain]
        ts += [c[1] for c in chain]
        for oacc in ('LL', 'MLL'):
            for nem in (1, 2, 'all'):
                res[f'ungroomed:{oacc}:{nem}'].append(O.ecf_from_chain_ungroomed(chain, 2., oacc, nem))
                res[f'softdrop:{oacc}:{nem}'].append(
                    O.ecf_from_chain_softdrop(chain, .1, 2., 0., oacc, nem))
            res[f'softdrop_bsd1:{oacc}:1'].append(O.ecf_from_chain_softdrop(chain, .1, 2., 1., oacc, 1))
    # z and theta are recomputed from rotated 3-vectors in the reference:
    # equal to the generated (z, theta) up to rounding in the geometry
    np.testing.assert_allclose(np.array(zs), g['z'], rtol=1e-12)
    np.testing.assert_allclose(np.array(ts), g["theta"], rtol=1e-5, atol=1e-15)
    for k, v in res.items():
        np.testing.assert_allclose(np.array(v), g[k], rtol=2e-5, err_msg=k)


def test_lin_log_mixed(golden):
    g = golden("lin_log_mixed")
    assert_array_equal(O.lin_log_mixed_list(1e-10, 1., 12), g['a'])
    assert_array_equal(O.lin_log_mixed_list(1e-15, 1., 501), g['b'])


@pytest.mark.parametrize("method", ['lin', 'log'])
def test_integrator_2d_and_grid_callers(golden, method):
    """integrator_2d and the numerical-radiator generators, restated with the oracle's pieces"""
    g = golden(f"callers_{method}")
    x, y, w = g['pre'], g['crit'][:, 1], g['i2d_weights']
    bins = O.make_edges_2d(12, [x, y], method)
    for name, kw in (('last', dict(last_bc=[0., 'plus'])), ('lastminus', dict(last_bc=[1., 'minus'])),
                     ('first', dict(first_bc=0.))):
        t = f'i2d_{name}:'
        assert_array_equal(bins[0], g[t + 'bins0'])
        assert_array_equal(bins[1], g[t + 'bins1'])
        r = O.integrate_hist_2d(x, y, w, bins, 2.5, **kw)
        for k in ('density', 'density_err', 'integral', 'integral_err', 'ext', 'ext_err'):
            assert_array_equal(r[k], g[t + k])
    # gen_numerical_radiator: bins from the whole (N,2) sample array, observable theta or C_ungroomed
    for jet, fc, acc in (('quark', True, 'LL'), ('gluon', False, 'MLL')):
        tag = f"{jet}_{'fc' if fc else 'rc'}:"
        for em, s, jac, area in (('crit', g['crit'], g['jac_crit'], float(g['area_crit'])),
                                 ('sub', g['sub'], g['jac_sub'], float(g['area_sub']))):
            z, th = s[:, 0], s[:, 1]
            obs = th if em == 'crit' else O.ecf_ungroomed(z, th, 2., acc)
            edges, mids = O.make_edges(30, s, method)
            assert_array_equal(edges, g[tag + em + ':bins'])
            r = O.integrate_hist(obs, O.weight_radiator(z, th, jet, fc, acc) * jac, edges, area, last_bc=[0., 'plus'])
            assert_array_equal(r['density'], g[tag + em + ':density'])
            assert_array_equal(r['integral'], g[tag + em + ':integral'])
            assert_array_equal(r['integral_err'], g[tag + em + ':integral error'])
        # gen_pre_num_rad
        z_em, th_c = g['pre'], g['crit'][:, 1]
        b2 = O.make_edges_2d(10, [z_em, th_c], method)
        wts = O.weight_radiator(z_em, th_c, jet, fc, acc)
        if method == 'lin':
            area, jacs = float(g['z_cut']), np.ones(len(th_c))
        else:
            area, jacs = np.log(1. / float(g['epsilon'])) * np.log(1. / float(g['epsilon'])), z_em * th_c
        r = O.integrate_hist_2d(z_em, th_c, wts * jacs, b2, area, last_bc=[0., 'plus'])
        assert_array_equal(np.array(b2), g[tag + 'pre:bins'])
        assert_array_equal(r['density'], g[tag + 'pre:density'])
        assert_array_equal(r['ext'], g[tag + 'pre:integral'])
        assert_array_equal(r['ext_err'], g[tag + 'pre:integral error'])
        # gen_crit_sub_num_rad loop
        grid = O.lin_log_mixed_list(float(g['epsilon']) if method == 'log' else 1e-8, 1., 10)
        s, jac, area0 = g['sub'], g['jac_sub'], float(g['area_sub'])
        xs_all, rads_all = [], []
        for thc in grid:
            th_em = s[:, 1] * thc
            obs = O.ecf_ungroomed(s[:, 0], th_em, 2., acc)
            edges, _ = O.make_edges(10, obs, method, thc ** 2. * 1e-15)
            area, jj = (area0 * thc, jac) if method == 'lin' else (area0, jac * thc)
            r = O.integrate_hist(obs, O.weight_radiator(s[:, 0], th_em, jet, fc, acc) * jj, edges, area,
                                 last_bc=[0., 'plus'])
            xs_all.append(edges)
            rads_all.append(r['integral'])
        assert_array_equal(np.array(xs_all).flatten(), g[tag + 'critsub:xs'])
        assert_array_equal(np.array(rads_all).flatten(), g[tag + 'critsub:integral'])


def test_vals_to_pdf(golden):
    g = golden("vals_to_pdf")
    xs, pdf = O.vals_to_pdf(g['vals'], 50, weights=g['weights'], log_cutoff=1e-10)
    assert_array_equal(xs, g['xs'])
    assert_array_equal(pdf, g['pdf'])


def test_loop_samplers_equal_vectorised_samplers():
    """the Python-loop restatement timed by bench.py's CPU baseline draws exactly the vectorised samples"""
    n = 500
    s, j = O.sample_loop_critical('log', n, .1, 1e-10, seed=5)
    s2, j2, _ = O.sample_critical('log', O.mt_uniforms(5, 2 * n).reshape(n, 2), .1, 1e-10)
    assert_array_equal(s, s2)
    assert_array_equal(j, j2)
    s, j = O.sample_loop_ungroomed('lin', n, seed=6, radius=.7)
    s2, j2, _ = O.sample_ungroomed('lin', O.mt_uniforms(6, 2 * n).reshape(n, 2), radius=.7)
    assert_array_equal(s, s2)


def test_bin_index_is_numpy_histogram():
    rng = np.random.RandomState(3)
    edges = np.logspace(-11, np.log10(.5), 100)
    obs = np.exp(rng.uniform(np.log(1e-13), 0., 20000))
    obs[:5] = [edges[0], edges[-1], edges[17], np.nan, 0.]
    idx = O.bin_index(obs, edges)
    n, _ = np.histogram(obs, edges)
    assert_array_equal(np.bincount(idx[idx >= 0], minlength=99), n)
    assert idx[0] == 0 and idx[1] == 98 and idx[2] == 17 and idx[3] == -1 and idx[4] == -1


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    r = O.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    r = O.philox4x32_10(f, f, f, f, f, f)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = O.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_argument_checks():
    with pytest.raises(AssertionError):
        O.sample_critical('log', np.zeros((1, 2)), .1, 0.)
    with pytest.raises(AssertionError):
        O.sample_critical('lin', np.zeros((1, 2)), .6)
    with pytest.raises(AssertionError):
        O.sample_ungroomed('cubic', np.zeros((1, 2)))
    with pytest.raises(AssertionError):
        O.sample_ungroomed('lin', np.zeros((1, 2)), radius=0.)


@pytest.mark.parametrize("jet,acc", [('quark', 'LL'), ('quark', 'MLL'), ('gluon', 'LL'), ('gluon', 'MLL')])
def test_shower_rss_replay(golden, jet, acc):
    """RSS grooming of the reference's own (z, theta) chains: bit for bit getECFs_rss (partonshower_utils.py:501-674)
    for every setting in oracle/make_golden.py::RSS_CASES"""
    chains_file, g = golden(f"shower_{jet}_{acc}"), golden(f"shower_rss_{jet}_{acc}")
    assert_array_equal(chains_file['lens'], g['lens'])            # same seeds -> same jets
    ends = np.cumsum(chains_file['lens'])
    chains = [list(zip(chains_file['z'][e - n:e], chains_file['theta'][e - n:e])) for e, n in zip(ends, chains_file['lens'])]
    keys = [k for k in g if k.startswith('rss:')]
    assert len(keys) == 24
    for k in keys:
        _, z_cut, f, oacc, etype, nem, few = k.split(':')
        nem = nem if nem == 'all' else int(nem)
        got = [O.ecf_from_chain_rss(c, float(z_cut), 2., f=float(f), obs_acc=oacc, emission_type=etype,
                                    n_emissions=nem, few_pres=bool(int(few))) for c in chains]
        assert_array_equal(np.array(got), g[k], err_msg=k)
    got = [O.ecf_from_chain_rss(c, 0, 2., obs_acc='LL', n_emissions=2) for c in chains]
    assert_array_equal(np.array(got), g['rss_zcut0:LL:2'])
    # the settings differ from one another (the vectors are not degenerate)
    assert not np.array_equal(g['rss:0.1:1.0:LL:precritsub:1:1'], g['rss:0.1:0.75:LL:precritsub:1:1'])
    assert not np.array_equal(g['rss:0.1:1.0:LL:precritsub:1:1'], g['rss:0.2:1.0:LL:precritsub:1:1'])


def _cdf_cases():
    """the synthetic cdfs of oracle/make_golden.py::CDF_CASES, restated (the generator script imports the reference
    and cannot be imported on the GPU box)"""
    return {
        'square': (lambda x: np.clip(x, 0, 1) ** 2, [0, 1], {}),
        'sudakov_pos_domain': (lambda x: np.exp(-np.log(1. / x) ** 2 * .3), [1e-8, 1.], {}),
        'neg_domain': (lambda x: .5 * (1 + np.tanh(x)), [-4., 4.], {}),
        'starts_at_one': (lambda x: np.where(x == 0, 1., .5 + .5 * x), [0, 1], {}),
        'negligible_then_monotone': (lambda x: np.where(x < .01, 1e-30 * (1 + np.sin(1e4 * x)), (x - .01) / .99),
                                     [0, 1], {}),
        'turnpoint': (lambda x: np.where(x < .6, np.minimum(x / .4, 1.), 1. - (x - .6)), [1e-6, 1.],
                      dict(catch_turnpoint=True)),
        'wiggle_forced': (lambda x: np.clip(x + .02 * np.sin(60 * x), 0, 1), [0, 1], dict(force_monotone=True)),
        'nan_head_forced': (lambda x: np.where(x < 1e-4, np.nan, np.clip(x + .05 * np.sin(25 * x), 0, 1)), [0, 1],
                            dict(force_monotone=True)),
    }


def test_samples_from_cdf_tabulated_branches(golden):
    """samples_from_cdf (montecarlo_utils.py:93-375) with pynverse refusing the cdf: every tabulated fall-back of
    the reference, bit for bit on the reference's np.random stream"""
    g = golden("samples_from_cdf")
    for name, (cdf, domain, kw) in _cdf_cases().items():
        rs = np.random.RandomState(int(g[name + ':seed']))
        with np.errstate(all='ignore'):
            samples, weights = O.samples_from_cdf_tabulated(cdf, 1500, domain, rs.rand, **kw)
        assert_array_equal(samples, g[name + ':samples'], err_msg=name)
        assert_array_equal(weights, g[name + ':weights'], err_msg=name)
    assert np.all(g['starts_at_one:samples'] == 0.)                      # zero bin
    assert g['turnpoint:samples'].max() <= .6 and g['negligible_then_monotone:samples'].min() >= .01
    for name in ('wiggle_forced', 'turnpoint'):
        assert bool(g[name + ':raises_without_flag'])
        cdf, domain, _ = _cdf_cases()[name]
        with pytest.raises(ValueError):
            O.samples_from_cdf_tabulated(cdf, 10, domain, np.random.RandomState(0).rand)


@pytest.mark.parametrize("method", ['log', 'lin'])
def test_process_and_ewoc_replay(golden, method):
    """MonteCarloProcess (hit-or-miss box sampling, efficiency-scaled area), the e+e- -> hadrons LO client, its
    pairwise observables and the EWOC built on them: the oracle on the reference's own uniform sequence, bit for bit.
    ref: montecarlo/process.py:109-191, montecarlo/ewoc.py:11-75, examples/processes/ee2hadrons_LO.py"""
    g = golden(f"process_ee2h_{method}")
    pre = O.ee2hadrons_lo_prefactor(1000.)
    assert pre == float(g['prefactor'])
    bounds = [(0, 1), (0, 1)]
    samples, weights, rejected = O.process_hit_or_miss(
        g['uniforms'], 3000, method, 1e-8, bounds, O.ee2h_in_phasespace,
        lambda a, b: O.ee2h_differential_xsec(a, b, pre))
    assert_array_equal(samples, g['samples'])
    assert_array_equal(weights, g['jacobians'])
    assert rejected == int(g['num_rejected']) and 2 * (3000 + rejected) == len(g['uniforms'])
    assert O.process_area(bounds, 3000, rejected) == float(g['area'])
    for name in ('costheta', 'm2', 'theta', 'z_i', 'z_j'):
        assert_array_equal(O.ee2h_pair_observable(samples, name), g[f'obs:{name}'], err_msg=name)
        assert_array_equal(np.tile(weights, 3), g[f'obs:{name}:weights'])
    assert_array_equal(O.ee2h_pair_observable(samples, 'gecf', kappa=1.5), g['obs:gecf1.5'])
    zi, zj = O.ee2h_pair_observable(samples, 'z_i'), O.ee2h_pair_observable(samples, 'z_j')
    area = float(g['area'])
    for name, spacing, ew in (('costheta', 'lin', (1, 1)), ('m2', 'lin', (1, 1)), ('m2', 'log', (2, 1)),
                              ('theta', 'log', (1, 1))):
        tag = f'ewoc:{name}:{spacing}:{ew[0]}{ew[1]}'
        obs = O.ee2h_pair_observable(samples, name)
        w = O.ewoc_weights(zi, zj, np.tile(weights, 3), ew)
        assert_array_equal(w, g[tag + ':weights'])
        edges = O.ewoc_bins(obs, 40, spacing)
        assert_array_equal(edges, g[tag + ':bins'])
        h, h2, c = O.histograms(obs, w, edges)
        dens, err = O.density_from_hist(h, h2, c, edges, area, len(obs))
        assert_array_equal(dens, g[tag + ':density'])
        assert_array_equal(err, g[tag + ':density_err'])
    obs = O.ee2h_pair_observable(samples, 'costheta')
    edges = O.ewoc_bins(obs, 25, 'lin', minval=-1., maxval=1.)
    assert_array_equal(edges, g['ewoc:range:bins'])
    h, h2, c = O.histograms(obs, O.ewoc_weights(zi, zj, np.tile(weights, 3)), edges)
    assert_array_equal(O.density_from_hist(h, h2, c, edges, area, len(obs))[0], g['ewoc:range:density'])
