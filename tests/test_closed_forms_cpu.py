"""Closed-form truth curves (critical Sudakov factor, Soft Drop radiators) against the reference's own functions
evaluated on observable grids (oracle/make_golden.py::closed_forms)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal


def test_critical_sudakov_closed_form(golden):
    from jetmontecarlo_b200.analytics.sudakov_factors.fixedcoupling import critSudakov_fc_LL
    g = golden("closed_forms")
    cs = g['c']
    n = 0
    for k in g:
        if not k.startswith('critSudakov:'):
            continue
        _, jet, z_c, beta, f = k.split(':')
        got = np.array(critSudakov_fc_LL(cs, float(z_c), float(beta), jet, f=float(f)), dtype=float)
        assert_array_equal(got, g[k], err_msg=k)
        n += 1
    assert n == 36
    # a cumulative distribution: rises to 1 at the end point C_max = 1/2 - z_c and stays there
    s = np.array(critSudakov_fc_LL(cs, .1, 2., 'quark'), dtype=float)
    assert np.all(np.diff(s) >= 0) and s[cs > .4].min() == 1. and 0 < s[0] < .5
    # known answers of SURVEY 8c
    assert_allclose(float(critSudakov_fc_LL(.01, .1, 2, 'quark')), 0.8865006864028651, rtol=1e-13)
    assert_allclose(float(critSudakov_fc_LL(.01, .1, 2, 'gluon')), 0.7661845529452946, rtol=1e-13)


def test_soft_drop_closed_forms(golden):
    from jetmontecarlo_b200.analytics.radiators import soft_drop as SD
    g = golden("closed_forms")
    cs = g['c']
    n = 0
    with np.errstate(all='ignore'):
        for k in g:
            parts = k.split(':')
            if parts[0] == 'sd_rad_fc':
                got = SD.softdrop_rad_fc(cs, float(parts[2]), float(parts[3]), jet_type=parts[1])
            elif parts[0] == 'sd_radprime_fc':
                got = SD.softdrop_radprime_fc(cs, float(parts[2]), float(parts[3]), jet_type=parts[1])
            elif parts[0] == 'sd_sudakov_fc_LL':
                got = SD.softdrop_sudakov_fc(cs, float(parts[2]), float(parts[3]), jet_type=parts[1])
            elif parts[0] == 'sd_sudakov_fc_ME':
                got = SD.softdrop_sudakov_fc(cs, float(parts[2]), float(parts[3]), jet_type=parts[1], acc='MLL')
            elif parts[0] == 'sd_rad_rc':
                got = SD.softdrop_rad_rc(cs, float(parts[2]), float(parts[3]), beta_sd=float(parts[4]), jet_type=parts[1])
                assert_array_equal(np.isnan(got), np.isnan(g[k]), err_msg=k)
                assert_allclose(got, g[k], rtol=1e-12, atol=1e-13, equal_nan=True, err_msg=k)     # regrouped products
                n += 1
                continue
            else:
                continue
            assert_array_equal(got, g[k], err_msg=k)
            n += 1
    assert n == 18 * 4 + 36
    assert SD.softdrop_rad_rc(cs, .1, 1.) == 0
    with pytest.raises(NotImplementedError):
        SD.softdrop_rad_rc(cs, .1, .5)
    # the fixed-coupling Sudakov factor is a cdf in c that reaches 1 at c = 1/2
    s = SD.softdrop_sudakov_fc(np.array([1e-6, 1e-3, .2, .5]), .1, 2.)
    assert np.all(np.diff(s) > 0) and s[-1] == 1.
    # (exactly at c = z_cut neither region's mask is set in the reference, :17-22: the radiator is 0 there)
    assert SD.softdrop_rad_fc(np.array([.1]), .1, 2.)[0] == 0.


def test_fixed_coupling_radiators_beyond_ll(golden):
    """ref: analytics/radiators/fixedcoupling.py:20-57, 110-160; running_coupling.py:298-312"""
    from jetmontecarlo_b200.analytics.radiators import fixedcoupling as FC
    from jetmontecarlo_b200.analytics.radiators import running_coupling as RC
    g = golden("closed_forms_fc")
    thetas, cs, zps = g['theta'], g['c'], g['z_pre']
    n = 0
    with np.errstate(all='ignore'):
        for jet in ('quark', 'gluon'):
            for z_c in (.05, .1, .2):
                assert_array_equal(FC.critRadPrimeAnalytic_fc_LL(thetas, z_c, jet), g[f'critRadPrime_fc_LL:{jet}:{z_c}'])
                assert_array_equal(FC.critRadAnalytic_fc(thetas, z_c, jet), g[f'critRad_fc:{jet}:{z_c}'])
                assert_array_equal(FC.critRadPrimeAnalytic_fc(thetas, z_c, jet), g[f'critRadPrime_fc:{jet}:{z_c}'])
                for th in (.3, 1e-2):
                    got = RC.preRadAnalytic_nofreeze(zps * z_c / .1, th, z_c, jet_type=jet)
                    assert_array_equal(got, g[f'preRad_nofreeze:{jet}:{z_c}:{th}'])
                    n += 1
            for beta in (1.5, 2., 3.):
                assert_array_equal(np.array(FC.subRadAnalytic_fc(cs, beta, jet), dtype=float), g[f'subRad_fc:{jet}:{beta}'])
                assert_array_equal(FC.subRadPrimeAnalytic_fc(cs, beta, jet), g[f'subRadPrime_fc:{jet}:{beta}'])
    assert n == 12
    # the reference's wrapper ignores its jet type: gluon == quark
    assert_array_equal(g['preRad_nofreeze:gluon:0.1:0.3'], g['preRad_nofreeze:quark:0.1:0.3'])
    with pytest.raises(AssertionError):
        FC.critRadAnalytic_fc(thetas, .1, 'photon')
