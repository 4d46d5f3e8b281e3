"""Fixed-coupling LL radiators and single-emission pdfs -- host mirror of
jetmontecarlo/analytics/radiators/fixedcoupling.py; the arithmetic runs in
libjmc_b200.so (csrc/jmc_physics.cuh).
"""
import numpy as np

from ... import _lib
from ..._arrayfn import eval_fn
from ..qcd_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..qcd_utils import _jet

_MAXR = {1: "max_radius"}


def critRadAnalytic_fc_LL(theta, z_c, jet_type='quark'):
    """ref: analytics/radiators/fixedcoupling.py:7-19"""
    return eval_fn(_lib.FN_CRIT_RAD_FC, [theta], jet=_jet(jet_type), z_cut=z_c)


def critPDFAnalytic_fc_LL(z, theta, z_c, jet_type='quark'):
    """ref: analytics/radiators/fixedcoupling.py:59-72"""
    return eval_fn(_lib.FN_CRIT_PDF_FC, [z, theta], jet=_jet(jet_type), z_cut=z_c)


def subRadAnalytic_fc_LL(C, beta, jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/fixedcoupling.py:77-90"""
    return eval_fn(_lib.FN_SUB_RAD_FC, [C, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


def subRadPrimeAnalytic_fc_LL(C, beta, jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/fixedcoupling.py:92-108"""
    return eval_fn(_lib.FN_SUB_RADPRIME_FC, [C, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


def subPDFAnalytic_fc_LL(C, beta, jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/fixedcoupling.py:162-176"""
    return eval_fn(_lib.FN_SUB_PDF_FC, [C, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


def preRadAnalytic_fc_LL(z_pre, theta_crit, z_cut, jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/fixedcoupling.py:181-192"""
    return eval_fn(_lib.FN_PRE_RAD_FC, [z_pre, theta_crit, maxRadius], jet=_jet(jet_type), z_cut=z_cut)


# ---- closed forms beyond the sampled path (host, on observable grids; the reference plots radiators against them) --
def critRadPrimeAnalytic_fc_LL(theta, z_c, jet_type='quark'):
    """ref: analytics/radiators/fixedcoupling.py:20-29"""
    from ..qcd_utils import CR, alpha_fc
    theta, z_c = theta + 1e-100, z_c + 1e-100
    return -2. * CR(jet_type) * np.log(2. * z_c) * alpha_fc / (np.pi * theta)


def _b_bar(z_c, jet_type):
    from ..qcd_utils import b_g_bar, b_q_bar, check_jet_type
    check_jet_type(jet_type)
    return b_q_bar(z_c) if jet_type == 'quark' else b_g_bar(z_c)


def critRadAnalytic_fc(theta, z_c, jet_type='quark'):
    """Critical radiator with the full hard-collinear coefficient b_bar(z_c).  ref: fixedcoupling.py:31-42"""
    from ..qcd_utils import R0, alpha_fc
    theta, z_c = theta + 1e-100, z_c + 1e-100
    return _b_bar(z_c, jet_type) * np.log(R0 / theta) * alpha_fc / np.pi


def critRadPrimeAnalytic_fc(theta, z_c, jet_type='quark'):
    """(minus) its derivative in theta.  ref: fixedcoupling.py:44-57"""
    from ..qcd_utils import alpha_fc
    theta, z_c = theta + 1e-100, z_c + 1e-100
    return _b_bar(z_c, jet_type) * alpha_fc / (np.pi * theta)


def subRadAnalytic_fc(C, beta=2., jet_type='quark'):
    """Subsequent radiator with full splitting functions (dilogarithms), maxRadius = 1.  ref: fixedcoupling.py:110-138"""
    from ..qcd_utils import CA, CF, N_F, TF, alpha_fc, check_jet_type, polylog_vec
    C = C + 1e-100
    check_jet_type(jet_type)
    if jet_type == 'quark':
        return CF * (9. - 18. * C + np.pi ** 2. - 12. * np.log(1. - C) * np.log(C) + 6. * np.log(C) ** 2.
                     + 9. * np.log(2. * C) - 12. * polylog_vec(2., 1. - C)) * alpha_fc / (6. * beta * np.pi)
    return (137. * CA - 288. * C * CA + 36. * C ** 2. * CA - 16. * C ** 3. * CA - 12. * CA * np.pi ** 2. - 88. * N_F * TF
            + 144. * C * N_F * TF - 72. * C ** 2. * N_F * TF + 32. * C ** 3. * TF + 72. * CA * np.log(C) ** 2.
            + 132. * CA * np.log(2. * C) - 48. * N_F * TF * np.log(2. * C)
            + 144. * CA * polylog_vec(2., C)) * alpha_fc / (72 * beta * np.pi)


def subRadPrimeAnalytic_fc(C, beta=2., jet_type='quark'):
    """(minus) its derivative in C.  ref: fixedcoupling.py:140-160"""
    from ..qcd_utils import CA, CF, N_F, TF, alpha_fc, check_jet_type
    C = C + 1e-100
    check_jet_type(jet_type)
    if jet_type == 'quark':
        return (-18. + 9. / C - 12. * np.log(1. - C) / C + 12. * np.log(C) / C) * CF * alpha_fc / (6 * beta * np.pi)
    return (-288. * CA + 132. * CA / C + 72. * C * CA - 48. * C ** 2. * CA + 144. * N_F * TF - 48 * N_F * TF / C
            - 144. * C * N_F * TF + 96 * C ** 2. * N_F * TF - 144. * CA * np.log(1. - C) / C
            + 144. * CA * np.log(C) / C) * alpha_fc / (72 * beta * np.pi)
