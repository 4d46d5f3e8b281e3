"""Closed-form Soft Drop radiators and Sudakov factors (Larkoski, Marzani, Soyez, Thaler; appendix A) -- the truth
curves for the Soft Drop groomed ECF distributions.  Host code on observable grids.  The reference module also
holds a plotting routine (and imports matplotlib helpers at module level); only the formulas are mirrored.
ref: analytics/radiators/soft_drop.py:15-194"""
import numpy as np
from scipy.special import gamma

from ..qcd_utils import CR, MU_NP, W, alpha1loop, alpha_fixed, b_i_sing, beta_0, check_jet_type, euler_constant


def softdrop_rad_fc(c, z_cut, beta, beta_sd=0, jet_type='quark'):
    """ref: analytics/radiators/soft_drop.py:15-23"""
    above = np.log(2. * c) ** 2. * (c > z_cut)
    below = (np.log(2. * z_cut) ** 2 - np.log(2. * z_cut) * np.log(z_cut / c) * 2.) * (c < z_cut)
    return CR(jet_type) * alpha_fixed / (beta * np.pi) * (above + below)


def softdrop_radprime_fc(c, z_cut, beta, beta_sd=0, jet_type='quark'):
    """ref: analytics/radiators/soft_drop.py:25-31"""
    return 2. * CR(jet_type) * alpha_fixed / (c * beta * np.pi) * (np.log(2. * c) * (c > z_cut)
                                                                   + np.log(2. * z_cut) * (c < z_cut))


def softdrop_sudakov_fc(c, z_cut, beta, beta_sd=0, jet_type='quark', acc='LL'):
    """exp(-R), times the multiple-emission factor exp(-gamma_E R') / Gamma(1 + R') unless acc == 'LL'.
    ref: analytics/radiators/soft_drop.py:33-43"""
    sudakov = np.exp(-softdrop_rad_fc(c, z_cut, beta, beta_sd, jet_type))
    if acc == 'LL':
        return sudakov
    log_slope = -c * softdrop_radprime_fc(c, z_cut, beta, beta_sd, jet_type)
    multiple_emissions = np.exp(-euler_constant * log_slope) / gamma(1. + log_slope)
    return sudakov * multiple_emissions


def softdrop_rad_rc(c, z_cut, beta, beta_sd=0, alpha=alpha_fixed, jet_type='quark'):
    """Running-coupling Soft Drop radiator in its four regions (above z_cut; down to the first non-perturbative
    bound; between the bounds; below).  Like the reference it evaluates every region and masks, so it is NaN where
    a logarithm's argument turns negative.  beta > 1 only: the reference's beta < 1 branches do not run (a call of
    a float at :77, an undefined name at :150), beta = 1 returns 0.
    ref: analytics/radiators/soft_drop.py:49-194"""
    check_jet_type(jet_type)
    if beta == 1:
        return 0
    if beta < 1:
        raise NotImplementedError("softdrop_rad_rc: the reference's beta < 1 branches raise (soft_drop.py:77,150)")
    Bi = b_i_sing(jet_type)
    pref = CR(jet_type) / (2. * np.pi * alpha * beta_0 ** 2.)
    L, Lc, Lmu = np.log(1. / c), np.log(1. / z_cut), np.log(1. / MU_NP)
    lam, lam_c, lam_mu = 2. * alpha * beta_0 * L, 2. * alpha * beta_0 * Lc, 2. * alpha * beta_0 * Lmu
    bsd1, bbsd, bm1 = 1. + beta_sd, beta + beta_sd, beta - 1.

    def F1(x):
        return (bsd1 * x - bbsd * Lmu + bm1 * Lc) ** 2. / (bm1 * bsd1 * bbsd)

    bound1 = z_cut ** ((1. - beta) / bsd1) * MU_NP ** (bbsd / bsd1)
    bound2 = MU_NP ** beta
    L1 = -2. * alpha * beta_0 * Bi * np.log(1. - lam / beta)
    frozen = CR(jet_type) * alpha1loop(MU_NP) / np.pi
    reg1 = pref * (L1 + W(1. - lam) / bm1 - beta * W(1. - lam / beta) / bm1)
    reg2 = pref * (L1 - W(1. - lam_c) / bsd1 - beta * W(1. - lam / beta) / bm1
                   + bbsd * W(1. - bsd1 / bbsd * lam - bm1 / bbsd * lam_c) / (bm1 * bsd1))
    reg3 = pref * (L1 - W(1. - lam_c) / bsd1 - beta * W(1. - lam / beta) / bm1
                   + (1. + np.log(1. - lam_mu)) * (bm1 * lam_c + bsd1 * lam - bbsd * lam_mu) / (bm1 * bsd1)
                   + bbsd * W(1. - lam_mu) / (bm1 * bsd1)) + frozen * F1(L)
    reg4 = pref * (-W(1. - lam_c) / bsd1 - beta_sd * W(1. - lam_mu) / bsd1
                   - 2. * alpha * beta_0 * Bi * np.log(1. - lam_mu)
                   - (1. + np.log(1. - lam_mu)) * (lam_c + beta_sd * lam_mu) / bsd1)\
        + frozen * (F1(beta * Lmu) + (L / beta - Lmu) * (2. * beta * Lc / bbsd + 2. * Bi
                                                         + beta_sd * (L + beta * Lmu) / bbsd))
    return reg1 * (z_cut <= c) + reg2 * (bound1 <= c) * (c < z_cut) + reg3 * (bound2 <= c) * (c < bound1)\
        + reg4 * (c < bound2)
