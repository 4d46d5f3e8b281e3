"""Fixed-coupling LL radiators and single-emission pdfs -- host mirror of
jetmontecarlo/analytics/radiators/fixedcoupling.py; the arithmetic runs in
libjmc_b200.so (csrc/jmc_physics.cuh).
"""
from ... import _lib
from ..._arrayfn import eval_fn
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
