"""Closed-form fixed-coupling Sudakov factor of a single critical emission -- the curve the reference's critical
tests plot their Monte Carlo cdfs against (examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py).  Host
code on observable grids (mpmath hypergeometric function), not on the per-sample path.
ref: analytics/sudakov_factors/fixedcoupling.py:11-51"""
import warnings

import numpy as np

from ..qcd_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..qcd_utils import CR, alpha_fc, hyp2f1_vec


def critSudakov_fc_LL(C, z_c, beta, jet_type, f=1., alpha=alpha_fc):
    """P(C_crit < C) for one critical emission at fixed coupling: the soft-collinear power law
    (C / (1/2 - f z_c))^eta, eta = 2 C_R alpha / (beta pi) ln(1 / (2 f z_c)), plus its hypergeometric corrections; 1
    above the end point C_max = 1/2 - f z_c.  For f != 1 the observable and z_c are rescaled first."""
    pref = (2. * CR(jet_type) * alpha) / (beta * np.pi)
    eta = pref * np.log(1. / (2. * f * z_c))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        x = C / (1. - (1. - f) * z_c)
        zc = f * z_c
        power_law = (x / (1. / 2. - zc)) ** eta
        scale = pref * (x / zc) ** eta / (eta * (eta - 1.))
        t1 = (zc / x) ** (eta - 1.) * hyp2f1_vec(1., 1. - eta, 2. - eta, -x / zc)
        t2 = (zc / x) ** eta * (eta - 1.) * np.log(1. + x / zc)
        t3 = -(zc / (1. / 2. - zc)) ** (eta - 1.) * hyp2f1_vec(1., 1. - eta, 2. - eta, 1. - 1. / (2. * zc))
        t4 = -(zc / (1. / 2. - zc)) ** eta * (eta - 1.) * np.log(1. / (2. * zc))
        sudakov = power_law + scale * (t1 + t2 + t3 + t4)
    c_max = (1. / 2. - f * z_c)
    return sudakov * (C < c_max) + 1. * (C > c_max)
