"""Running-coupling (MLL) radiators, their derivatives and single-emission
pdfs -- host mirror of jetmontecarlo/analytics/radiators/running_coupling.py.
Like the reference, every region is evaluated and masked, so the results are
NaN below the Landau scale; callers apply np.nan_to_num.
"""
import numpy as np

from ... import _lib
from ..._arrayfn import eval_fn
from ..qcd_utils import *  # noqa: F401,F403  (the reference's module re-exports the QCD constants the same way)
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


# ---- the regions the radiators are assembled from (the reference's analytic MLL curves use them piecewise) --------
def _piece(fn, x, alpha, jet_type, **kw):
    return eval_fn(fn, [x, alpha], jet=_jet(jet_type), **kw)


def subSC1(C, beta=2, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:11-29"""
    return _piece(_lib.FN_SUB_SC1, C, alpha, jet_type, beta=beta)


def subSC2(C, beta=2, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:30-56"""
    return _piece(_lib.FN_SUB_SC2, C, alpha, jet_type, beta=beta)


def subSC3(C, beta=2, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:57-72"""
    return _piece(_lib.FN_SUB_SC3, C, alpha, jet_type, beta=beta)


def subHC1(C, beta=2, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:77-95"""
    return _piece(_lib.FN_SUB_HC1, C, alpha, jet_type, beta=beta)


def subHC2(C, beta=2, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:97-117"""
    return _piece(_lib.FN_SUB_HC2, C, alpha, jet_type, beta=beta)


def critSC1(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:157-175"""
    return _piece(_lib.FN_CRIT_SC1, theta, alpha, jet_type, z_cut=z_c)


def critSC2(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:177-199"""
    return _piece(_lib.FN_CRIT_SC2, theta, alpha, jet_type, z_cut=z_c)


def critSC3(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:201-219"""
    return _piece(_lib.FN_CRIT_SC3, theta, alpha, jet_type, z_cut=z_c)


def critHC1(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:224-240"""
    return _piece(_lib.FN_CRIT_HC1, theta, alpha, jet_type, z_cut=z_c)


def critHC2(theta, z_c, jet_type='quark', alpha=alpha_fixed):
    """ref: analytics/radiators/running_coupling.py:242-258"""
    return _piece(_lib.FN_CRIT_HC2, theta, alpha, jet_type, z_cut=z_c)


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
