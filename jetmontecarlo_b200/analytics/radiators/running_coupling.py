"""Running-coupling (MLL) radiators, their derivatives and single-emission
pdfs -- host mirror of jetmontecarlo/analytics/radiators/running_coupling.py.
Like the reference, every region is evaluated and masked, so the results are
NaN below the Landau scale; callers apply np.nan_to_num.
"""
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
