"""Phase-space weight functions -- host mirror of jetmontecarlo/numerics/weights.py."""
import numpy as np

from .. import _lib
from .._arrayfn import eval_fn
from ..analytics.qcd_utils import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..analytics.radiators.running_coupling import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..analytics.radiators.fixedcoupling import *  # noqa: F401,F403  (re-exported as the reference's module does)
from ..analytics.qcd_utils import _jet


def radiatorWeight(z, theta, jet_type, fixedcoupling=True, acc='LL'):
    """splittingFn * alpha / (theta pi).  ref: numerics/weights.py:15-26"""
    assert acc in ['LL', 'LL_nonsing', 'MLL'], "Invalid accuracy."
    return eval_fn(_lib.FN_RADIATOR_WEIGHT, [z, theta], jet=_jet(jet_type), acc=_lib.ACCS[acc],
                   fixed_coupling=fixedcoupling)


def criticalEmissionWeight(z, theta, z_c, jet_type='quark', fixedcoupling=True):
    """ref: numerics/weights.py:31-41"""
    fn = _lib.FN_CRIT_PDF_FC if fixedcoupling else _lib.FN_CRIT_PDF_RC
    return eval_fn(fn, [z, theta], jet=_jet(jet_type), z_cut=z_c)


def precriticalEmissionWeight(z_pre, theta_crit, z_c, jet_type='quark', fixedcoupling=True):
    """ref: numerics/weights.py:43-58"""
    if not fixedcoupling:
        raise TypeError("Can only accept fixed coupling for now!")
    return eval_fn(_lib.FN_PRE_WEIGHT, [z_pre, theta_crit], jet=_jet(jet_type), z_cut=z_c)


def precriticalEmissionWeight_zerobin(z_pre, theta_crit, z_c, jet_type='quark', fixedcoupling=True,
                                      epsilon=None):
    """ref: examples/tests/comparison_tests/comparison_utils_probabilistic.py:82-98"""
    assert epsilon is not None and 0 < epsilon < 1, "Weight requires valid cutoff for zero bin."
    if not fixedcoupling:
        raise TypeError("Can only accept fixed coupling for now!")
    return eval_fn(_lib.FN_PRE_WEIGHT_ZEROBIN, [z_pre, theta_crit], jet=_jet(jet_type), z_cut=z_c,
                   epsilon=epsilon)


def twoEmissionWeight(critsample, subsample, z_c, beta, jet_type='quark', fixedcoupling=True):
    """Critical x subsequent weight on (N,2) sample arrays; keeps the reference's
    ``csub = z_s th_s^b th_s^b``.  ref: numerics/weights.py:63-99"""
    assert len(critsample) == len(subsample), \
        "Must have the same amount of critical and subsequent samples!"
    return eval_fn(_lib.FN_TWO_EMISSION,
                   [critsample[:, 0], critsample[:, 1], subsample[:, 0], subsample[:, 1]],
                   jet=_jet(jet_type), z_cut=z_c, beta=beta, fixed_coupling=fixedcoupling)
