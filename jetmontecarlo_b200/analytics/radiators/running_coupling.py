"""Running-coupling (MLL) radiators, their derivatives and single-emission
pdfs -- host mirror of jetmontecarlo/analytics/radiators/running_coupling.py.
Like the reference, every region is evaluated and masked, so the results are
NaN below the Landau scale; callers apply np.nan_to_num.
"""
import numpy as np

from ... import _lib
from ..._arrayfn import eval_fn
from ..qcd_utils import _jet, alpha_fixed

_MAXR = {1: "max_radius"}


def _default_alpha(alpha):
    if alpha != alpha_fixed:
        raise NotImplementedError("only the default alpha=alpha_fixed is on the accelerated path")


def critRadAnalytic(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:263-289"""
    _default_alpha(alpha)
    return eval_fn(_lib.FN_CRIT_RAD_RC, [theta], jet=_jet(jet_type), z_cut=z_c)


def critRadPrimeAnalytic(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:326-347"""
    _default_alpha(alpha)
    return eval_fn(_lib.FN_CRIT_RADPRIME_RC, [theta], jet=_jet(jet_type), z_cut=z_c)


def critPDFAnalytic(z, theta, z_c, jet_type='quark'):
    """ref: analytics/radiators/running_coupling.py:386-404"""
    return eval_fn(_lib.FN_CRIT_PDF_RC, [z, theta], jet=_jet(jet_type), z_cut=z_c)


def subRadAnalytic(C, beta=2, jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/running_coupling.py:119-148"""
    return eval_fn(_lib.FN_SUB_RAD_RC, [C, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


def subRadPrimeAnalytic(c, beta=2., jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/running_coupling.py:350-380"""
    return eval_fn(_lib.FN_SUB_RADPRIME_RC, [c, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


def subPDFAnalytic(c, beta=2., jet_type='quark', maxRadius=1.):
    """ref: analytics/radiators/running_coupling.py:409-422"""
    return eval_fn(_lib.FN_SUB_PDF_RC, [c, maxRadius], jet=_jet(jet_type), beta=beta, scalar_fields=_MAXR)


# ---- pre-critical radiator without frozen coupling (host closed form) ---------------------------------------------
def preSC_nofreeze(z_pre, theta_crit, z_cut, alpha=None, jet_type='quark'):
    """ref: analytics/radiators/running_coupling.py:298-307"""
    from ..qcd_utils import CR, Lambda, W, alpha_fixed, beta_0, check_jet_type
    alpha = alpha_fixed if alpha is None else alpha
    check_jet_type(jet_type)
    sc = W(1. + Lambda(z_cut, alpha)) - W(1. + Lambda(z_pre, alpha)) - W(1. + Lambda(z_cut * theta_crit, alpha))\
        + W(1. + Lambda(z_pre * theta_crit, alpha))
    return CR(jet_type) / (2. * alpha * beta_0 ** 2. * np.pi) * sc


def preRadAnalytic_nofreeze(z_pre, theta_crit, z_cut, alpha=None, jet_type='quark'):
    """ref: analytics/radiators/running_coupling.py:309-312 -- which passes alpha_fixed and 'quark' on whatever its
    own arguments say; reproduced."""
    return preSC_nofreeze(z_pre, theta_crit, z_cut, alpha=None, jet_type='quark')
