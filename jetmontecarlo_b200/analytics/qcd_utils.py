"""QCD constants, coupling and splitting functions -- host mirror of
jetmontecarlo/analytics/qcd_utils.py (same names, arguments and assertion
messages).  The per-sample functions (alpha_s, splittingFn and friends) run in
the CUDA kernels of libjmc_b200.so; only import-time constants are computed on
the host.
"""
import numpy as np

from .. import _lib
from .._arrayfn import eval_fn

# ---- jet types                                   ref: analytics/qcd_utils.py:16-28
valid_jet_types = ['quark', 'gluon']


def check_jet_type(jet_type):
    assert jet_type in valid_jet_types, \
        str(jet_type) + "is not a supported jet type."


# ---- group theory                                ref: analytics/qcd_utils.py:33-66
CF = 8. / 6.
CA = 3.
TF = 1. / 2.
N_F = 5.


def CR(jet_type):
    check_jet_type(jet_type)
    return CF if jet_type == 'quark' else CA


beta_0 = (33 - 2 * N_F) / (12 * np.pi)
B_Q = -3. / 4.
B_G = -11. / 12. + N_F / (6 * CA)


def b_i_sing(jet_type):
    check_jet_type(jet_type)
    return B_Q if jet_type == 'quark' else B_G


M_Z = 91.19
P_T = 3000.
euler_constant = 0.5772156649


def alpha1loop(mu):
    """One-loop coupling at mu [GeV] (import-time constants only; per-sample
    couplings are alpha_s).  ref: analytics/qcd_utils.py:102-105"""
    alpha_s_zmass = 0.118
    return alpha_s_zmass / (1 + 2 * alpha_s_zmass * beta_0 * np.log(mu / M_Z))


def alpha_em_1loop(mu):
    """One-loop electromagnetic coupling at mu [GeV].  ref: analytics/qcd_utils.py:108-112"""
    alpha_em_zmass = 1. / 127
    beta0_em = 1. / (3 * np.pi)
    return alpha_em_zmass / (1 + 2 * alpha_em_zmass * beta0_em * np.log(mu / M_Z))


def sum_square_quark_charges(energy):
    """Sum of the squared charges of the quarks that can be produced at ``energy`` [GeV], in units of e (u, d above
    10 MeV; s above 50 MeV; c above 2.5 GeV; b above 10 GeV; no top).  ref: analytics/qcd_utils.py:72-96"""
    thresholds = ((0.01, 1 / 3), (0.01, 2 / 3), (0.05, -1 / 3), (2.5, 2 / 3), (10.0, -1 / 3))
    square_charges = 0
    for threshold, charge in thresholds:
        if energy > threshold:
            square_charges += charge ** 2
    return square_charges


R0 = 1.
alpha_fixed = alpha1loop(P_T * R0)      # ref: qcd_utils.py:196
MU_NP = 1. / (P_T * R0)
LAMBDA_QCD = .3 / (P_T * R0)
TEN_MeV = .01 / (P_T * R0)
ONE_MeV = .001 / (P_T * R0)
alpha_fc = alpha_fixed


def _jet(jet_type):
    check_jet_type(jet_type)
    return _lib.JETS[jet_type]


def alpha_s(z, theta):
    """Coupling at the scale of a splitting, frozen below 1 GeV.
    ref: analytics/qcd_utils.py:115-119"""
    return eval_fn(_lib.FN_ALPHA_S, [z, theta])


def splittingFn(z, jet_type, accuracy):
    """ref: analytics/qcd_utils.py:160-190"""
    assert accuracy in ['LL', 'LL_nonsing', 'MLL'], "Invalid accuracy."
    return eval_fn(_lib.FN_SPLITTING, [z], jet=_jet(jet_type), acc=_lib.ACCS[accuracy])


def splitting(z, jet_type):
    """ref: analytics/qcd_utils.py:134-140"""
    return splittingFn(z, jet_type, 'LL_nonsing')


def splitting_bar(z, jet_type):
    """ref: analytics/qcd_utils.py:149-157"""
    return splittingFn(z, jet_type, 'MLL')


def q_splitting(z):
    return splitting(z, 'quark')


def g_splitting(z):
    return splitting(z, 'gluon')


def q_splitting_bar(z):
    return splitting_bar(z, 'quark')


def g_splitting_bar(z):
    return splitting_bar(z, 'gluon')


# ---- small closed-form helpers (host; used by the analytic truth curves, never per Monte Carlo sample) ---------
def W(x):
    """W(x) = x ln x.  ref: analytics/qcd_utils.py:211-216"""
    return x * np.log(x)


def Lambda(x, alpha):
    """ref: analytics/qcd_utils.py:218-221"""
    return 2. * alpha * beta_0 * np.log(x)


def b_q_bar(z_c):
    """ref: analytics/qcd_utils.py:223-227"""
    return CF * (-3. + 6. * z_c + 4. * np.log(2. - 2. * z_c)) / 2.


def b_g_bar(z_c):
    """ref: analytics/qcd_utils.py:229-236"""
    return ((-1. + 2. * z_c) * (-4. * N_F * TF * (1. + (-1. + z_c) * z_c) + CA * (11. + 2. * (-1. + z_c) * z_c)) / 6.
            + 2. * CA * np.log(2. - 2. * z_c))


def _mpmath_vec(name, nin):
    import mpmath
    return np.frompyfunc(getattr(mpmath, name), nin, 1)


def polylog_vec(*args):
    """np-friendly mpmath.polylog.  ref: analytics/qcd_utils.py:239"""
    return _mpmath_vec('polylog', 2)(*args)


def hyp2f1_vec(*args):
    """np-friendly mpmath.hyp2f1.  ref: analytics/qcd_utils.py:240"""
    return _mpmath_vec('hyp2f1', 4)(*args)
