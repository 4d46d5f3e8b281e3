"""CPU oracle for the JetMonteCarlo Monte Carlo integration hot path.

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of what the
reference (samcaf/JetMonteCarlo, read-only at /root/reference in the build
container) computes on the path

    sample emission phase space -> jacobian x weight -> weighted histogram
    -> density +- error -> cumulative integral +- error.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``jetmontecarlo_b200`` never does: it fails loudly when its CUDA library is
missing instead of falling back to this code.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference (through ``oracle/ref_shim.py``) and stores its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those vectors *bit for bit* (same NumPy ufuncs, same operation order)
and against the known-answer values listed in SURVEY.md 8(c).

Every function cites the reference lines it follows as ``ref: file:lines``
(paths relative to /root/reference/jetmontecarlo unless stated).

Third-party arithmetic the reference relies on and that is NOT under
/root/reference: ``numpy`` (ufuncs, ``np.histogram``, ``np.cumsum``; reference
pins numpy==1.23.1, this image has 2.3.x) and CPython's ``random`` (MT19937,
53-bit doubles).  ``np.histogram`` is restated per sample in ``bin_index`` and
checked against NumPy itself.  ``scipy.interpolate.interp1d`` (the installed
SciPy) is used as is by ``inverse_transform``.  ``pynverse`` 0.1.4.4
(``inversefunc``, utils/montecarlo_utils.py:6,141) is not installed: of
``samples_from_cdf`` only the tabulated route the reference takes when
``inversefunc`` refuses a cdf is restated (``samples_from_cdf_tabulated``) and
pinned; its root-finding route is UNPINNED and not restated.
"""
import math
import random as _pyrandom

import numpy as np

# --------------------------------------------------------------------------
# QCD constants                                  ref: analytics/qcd_utils.py:33-66,195-206
# --------------------------------------------------------------------------
CF = 8. / 6.
CA = 3.
TF = 1. / 2.
N_F = 5.
BETA_0 = (33 - 2 * N_F) / (12 * np.pi)
M_Z = 91.19
P_T = 3000.
R0 = 1.
ALPHA_MZ = 0.118
QUARK, GLUON = 'quark', 'gluon'
LL, LL_NONSING, MLL = 'LL', 'LL_nonsing', 'MLL'


def casimir(jet):
    """ref: analytics/qcd_utils.py:38-44"""
    if jet not in (QUARK, GLUON):
        raise AssertionError(str(jet) + "is not a supported jet type.")
    return CF if jet == QUARK else CA


def alpha_one_loop(mu):
    """One-loop coupling at scale mu [GeV].  ref: analytics/qcd_utils.py:102-105"""
    return ALPHA_MZ / (1 + 2 * ALPHA_MZ * BETA_0 * np.log(mu / M_Z))


ALPHA_FIXED = alpha_one_loop(P_T * R0)          # ref: qcd_utils.py:196
MU_NP = 1. / (P_T * R0)                         # ref: qcd_utils.py:198


def alpha_running(z, theta):
    """Coupling frozen below 1 GeV.  ref: analytics/qcd_utils.py:115-119"""
    return alpha_one_loop(np.maximum(P_T * z * theta, 1.))


# --------------------------------------------------------------------------
# Splitting functions                             ref: analytics/qcd_utils.py:124-190
# --------------------------------------------------------------------------
def _p_quark(z):
    return CF * ((1 + (1 - z) ** 2.) / z)


def _p_gluon(z):
    return CA * (2. * (1. - z) / z + z * (1. - z)
                 + (TF * N_F) * (z ** 2. + (1. - z) ** 2.) / CA)


def splitting(z, jet, acc):
    """LL: 2 C_R / z; LL_nonsing: P(z); MLL: P(z) + P(1-z).
    ref: analytics/qcd_utils.py:160-190"""
    if acc not in (LL, LL_NONSING, MLL):
        raise AssertionError("Invalid accuracy.")
    cr = casimir(jet)
    if acc == LL:
        return 2. * cr / z
    p = _p_quark if jet == QUARK else _p_gluon
    if acc == LL_NONSING:
        return p(z)
    return p(z) + p(1 - z)


def b_bar(zc, jet):
    """Integrated hard-collinear coefficient.  ref: analytics/qcd_utils.py:223-236"""
    if jet == QUARK:
        return CF * (-3. + 6. * zc + 4. * np.log(2. - 2. * zc)) / 2.
    return ((-1. + 2. * zc) * (-4. * N_F * TF * (1. + (-1. + zc) * zc)
                               + CA * (11. + 2. * (-1. + zc) * zc)) / 6.
            + 2. * CA * np.log(2. - 2. * zc))


def _xlogx(x):
    """W(x).  ref: analytics/qcd_utils.py:211-216"""
    return x * np.log(x)


def _lam(x, alpha):
    """Lambda(x, alpha) = 2 alpha beta0 ln x.  ref: analytics/qcd_utils.py:218-221"""
    return 2. * alpha * BETA_0 * np.log(x)


def _colour_norm(piece, jet, alpha, coeff=None):
    """(coeff*piece)*alpha/(pi*2*alpha*beta0) with the reference's operation
    order.  ref: e.g. analytics/radiators/running_coupling.py:24-29"""
    if coeff is None:
        coeff = 2 * CF if jet == QUARK else 2 * CA
    return (coeff * piece) * alpha / (np.pi * 2. * alpha * BETA_0)


# --------------------------------------------------------------------------
# Observables                                     ref: numerics/observables.py:15-114
# --------------------------------------------------------------------------
def ecf_groomed(z_crit, theta_crit, z_cut, beta, z_pre=0., f=1., acc=LL):
    """ref: numerics/observables.py:15-79"""
    if acc not in (LL, MLL):
        raise AssertionError("Accuracy must be LL or MLL!")
    zc_left = z_cut - z_pre
    zmin = np.minimum(z_crit, 1. - z_crit)
    if acc == LL:
        return ((zmin - f * zc_left) * (1. - (1. - f) * zc_left)
                * theta_crit ** beta) / (1. - z_cut) ** 2.
    return ((zmin - f * zc_left) * (1. - zmin - (1. - f) * zc_left)
            * theta_crit ** beta) / (1. - z_cut) ** 2.


def ecf_ungroomed(z, theta, beta, acc=LL):
    """ref: numerics/observables.py:82-109"""
    if acc not in (LL, MLL):
        raise AssertionError("Accuracy must be LL or MLL")
    if acc == LL:
        return z * theta ** beta
    return z * (1. - z) * theta ** beta


def ecf_ungroomed_max(beta, radius, acc):
    """ref: numerics/observables.py:111-114"""
    return radius ** beta / 2. if acc == LL else radius ** beta / 4.


# --------------------------------------------------------------------------
# Fixed-coupling analytics                        ref: analytics/radiators/fixedcoupling.py
# --------------------------------------------------------------------------
_TINY = 1e-100


def pdf_crit_fc(z, theta, z_c, jet=QUARK):
    """ref: analytics/radiators/fixedcoupling.py:59-72"""
    theta = theta + _TINY
    z_c = z_c + _TINY
    cr = casimir(jet)
    return (2. * cr * ALPHA_FIXED / (np.pi * theta * z)
            * np.exp(np.log(1. / (2. * z_c)) * np.log(theta / R0)
                     * 2. * cr * ALPHA_FIXED / np.pi)
            * (z_c < z) * (z < 1. / 2.))


def rad_crit_fc(theta, z_c, jet=QUARK):
    """ref: analytics/radiators/fixedcoupling.py:7-19"""
    theta = theta + _TINY
    z_c = z_c + _TINY
    return (-2. * casimir(jet) * np.log(R0 / theta)
            * np.log(2. * z_c) * ALPHA_FIXED / np.pi)


def rad_sub_fc(C, beta, jet=QUARK, max_radius=1.):
    """ref: analytics/radiators/fixedcoupling.py:77-90"""
    C = C + _TINY
    rad = (casimir(jet) * ALPHA_FIXED / (beta * np.pi)
           * np.log(2. * C / max_radius ** beta) ** 2.)
    return rad * (C < max_radius ** beta / 2.)


def radprime_sub_fc(C, beta, jet=QUARK, max_radius=1.):
    """ref: analytics/radiators/fixedcoupling.py:92-108"""
    C = C + _TINY
    d = (2. * casimir(jet) * ALPHA_FIXED / (beta * np.pi)
         * np.log(2. * C / max_radius ** beta) / C)
    return d * (C < max_radius ** beta / 2.)


def pdf_sub_fc(C, beta, jet=QUARK, max_radius=1.):
    """ref: analytics/radiators/fixedcoupling.py:162-176 (note the second +1e-100
    applied inside the two callees, as in the reference)."""
    C = C + _TINY
    sud = np.exp(-rad_sub_fc(C, beta, jet, max_radius))
    return sud * (-radprime_sub_fc(C, beta, jet, max_radius))


def rad_pre_fc(z_pre, theta_crit, z_cut, jet=QUARK, max_radius=1.):
    """ref: analytics/radiators/fixedcoupling.py:181-192"""
    rad = (2 * casimir(jet) * ALPHA_FIXED / np.pi
           * np.log(z_pre / z_cut) * np.log(theta_crit / max_radius))
    return (rad * (0. < z_pre) * (z_pre < z_cut)
            * (0. < theta_crit) * (theta_crit < max_radius))


def sudakov_crit_fc(C, z_c, beta, jet, f=1., alpha=None):
    """Closed-form cdf of one fixed-coupling critical emission in C = (z - f z_c) theta^beta (mpmath 2F1).
    ref: analytics/sudakov_factors/fixedcoupling.py:11-51 (critSudakov_fc_LL)"""
    import mpmath
    alpha = ALPHA_FIXED if alpha is None else alpha
    cr = casimir(jet)
    eta = (2. * cr * alpha) / (beta * np.pi) * np.log(1. / (2. * f * z_c))
    C = np.atleast_1d(np.asarray(C, dtype=float))
    Cs, zc = C / (1. - (1. - f) * z_c), f * z_c
    h = np.frompyfunc(lambda x: float(mpmath.hyp2f1(1., 1. - eta, 2. - eta, x)), 1, 1)
    with np.errstate(all='ignore'):
        simple = (Cs / (1. / 2. - zc)) ** eta
        prefac = (2. * cr * alpha / (beta * np.pi)) * (Cs / zc) ** eta / (eta * (eta - 1.))
        c1 = (zc / Cs) ** (eta - 1.) * np.asarray(h(-Cs / zc), dtype=float)
        c2 = (zc / Cs) ** eta * (eta - 1.) * np.log(1. + Cs / zc)
        c3 = -(zc / (1. / 2. - zc)) ** (eta - 1.) * float(mpmath.hyp2f1(1., 1. - eta, 2. - eta, 1. - 1. / (2. * zc)))
        c4 = -(zc / (1. / 2. - zc)) ** eta * (eta - 1.) * np.log(1. / (2. * zc))
        full = simple + prefac * (c1 + c2 + c3 + c4)
    c_max = 1. / 2. - f * z_c
    return full * (C < c_max) + 1. * (C > c_max)


# --------------------------------------------------------------------------
# Running-coupling analytics                      ref: analytics/radiators/running_coupling.py
# Every region is evaluated and then masked, exactly as the reference does, so
# log(negative) -> NaN and NaN * 0 -> NaN below the Landau scale.
# --------------------------------------------------------------------------
def _sub_sc1(C, beta, jet, alpha):
    """ref: running_coupling.py:11-29"""
    sc = beta / (2. * alpha * BETA_0 * (beta - 1.)) * (
        _xlogx(1. + _lam(1. / 2., alpha)) * (1. - 1. / beta)
        - _xlogx(1. + _lam(C / 2. ** (beta - 1.), alpha) / beta)
        + _xlogx(1. + _lam(C, alpha)) / beta)
    return _colour_norm(sc, jet, alpha)


def _sub_sc2(C, beta, jet, alpha):
    """ref: running_coupling.py:31-56"""
    pert = beta / (2. * alpha * BETA_0 * (beta - 1.)) * (
        _xlogx(1. + _lam(1. / 2., alpha) * (1. - 1. / beta)
               + _lam(MU_NP, alpha) / beta)
        - _xlogx(1. + _lam(1. / 2., alpha) * (1. - 1. / beta)
                 + _lam(C, alpha) / beta)
        + _lam(C / MU_NP, alpha) / beta
        * (1 + np.log(1 + _lam(MU_NP, alpha))))
    nonpert = 2. * alpha * BETA_0 * (np.log(MU_NP / C)) ** 2. / (2. * (beta - 1.))
    return _colour_norm(pert + nonpert, jet, alpha)


def _sub_sc3(C, beta, jet, alpha):
    """ref: running_coupling.py:58-72"""
    sc = (np.log(2. * C) ** 2. / beta - beta * np.log(2. * MU_NP) ** 2.
          ) * 2. * alpha * BETA_0 / 2.
    return _colour_norm(sc, jet, alpha)


def _sub_hc1(C, beta, jet, alpha):
    """ref: running_coupling.py:77-95"""
    hc = (np.log(1 + _lam(1. / 2., alpha))
          - np.log(1 + _lam(C / 2. ** (beta - 1.), alpha) / beta))
    return _colour_norm(hc, jet, alpha, coeff=b_bar(C, jet))


def _sub_hc2(C, beta, jet, alpha):
    """ref: running_coupling.py:97-117"""
    hc = (2. * alpha * BETA_0 * np.log(MU_NP ** beta * 2. ** (beta - 1) / C)
          / ((1. + _lam(MU_NP, alpha)) * beta))
    return _colour_norm(hc, jet, alpha, coeff=b_bar(C, jet))


def rad_sub_rc(C, beta=2, jet=QUARK, max_radius=1.):
    """ref: running_coupling.py:119-148"""
    with np.errstate(all='ignore'):
        C = C / max_radius ** beta
        alpha = alpha_one_loop(np.maximum(P_T * max_radius, 1.))
        knee = MU_NP ** beta * 2 ** (beta - 1.)
        soft = ((C > MU_NP) * _sub_sc1(C, beta, jet, alpha)
                + (knee < C) * (C < MU_NP)
                * (_sub_sc1(MU_NP, beta, jet, alpha) + _sub_sc2(C, beta, jet, alpha))
                + (C < knee)
                * (_sub_sc1(MU_NP, beta, jet, alpha)
                   + _sub_sc2(knee, beta, jet, alpha)
                   + _sub_sc3(C, beta, jet, alpha)))
        hard = ((C > knee) * _sub_hc1(C, beta, jet, alpha)
                + (C < knee)
                * (_sub_hc1(knee, beta, jet, alpha) + _sub_hc2(C, beta, jet, alpha)))
        return (soft + hard) * (C < 1. / 2.)


def _crit_sc1(theta, z_c, jet, alpha):
    """ref: running_coupling.py:157-175"""
    sc = (_xlogx(1. + _lam(1. / 2., alpha))
          - _xlogx(1. + _lam(z_c, alpha))
          - _xlogx(1. + _lam(theta / 2., alpha))
          + _xlogx(1. + _lam(z_c * theta, alpha))) / (2. * alpha * BETA_0)
    return _colour_norm(sc, jet, alpha)


def _crit_sc2(theta, z_c, jet, alpha):
    """ref: running_coupling.py:177-199"""
    pert = (_xlogx(1. + _lam(MU_NP / (2. * z_c), alpha))
            - _xlogx(1 + _lam(theta / 2., alpha))
            + _lam(theta * z_c / MU_NP, alpha)
            * (1. + np.log(1 + _lam(MU_NP, alpha)))) / (2. * alpha * BETA_0)
    nonpert = ((2. * alpha * BETA_0) * np.log(theta * z_c / MU_NP) ** 2.
               / (2. * (1. + _lam(MU_NP, alpha))))
    return _colour_norm(pert + nonpert, jet, alpha)


def _crit_sc3(theta, z_c, jet, alpha):
    """ref: running_coupling.py:201-219"""
    sc = (-(2. * alpha * BETA_0) * np.log(2. * z_c) * np.log(2. * MU_NP / theta)
          / (1. + _lam(MU_NP, alpha)))
    return _colour_norm(sc, jet, alpha)


def _crit_hc1(theta, z_c, jet, alpha):
    """ref: running_coupling.py:224-240"""
    hc = -np.log(1. + _lam(theta, alpha))
    return _colour_norm(hc, jet, alpha, coeff=b_bar(z_c, jet))


def _crit_hc2(theta, z_c, jet, alpha):
    """ref: running_coupling.py:242-258"""
    hc = _lam(MU_NP * 2. / theta, alpha) / (1. + _lam(MU_NP, alpha))
    return _colour_norm(hc, jet, alpha, coeff=b_bar(z_c, jet))


def rad_crit_rc(theta, z_c, jet=QUARK, alpha=ALPHA_FIXED):
    """ref: running_coupling.py:263-289.  The third soft region calls critSC2 /
    critSC3 with the *default* alpha (:276-277); kept."""
    with np.errstate(all='ignore'):
        soft = ((theta > MU_NP / z_c) * _crit_sc1(theta, z_c, jet, alpha)
                + (2. * MU_NP < theta) * (theta < MU_NP / z_c)
                * (_crit_sc1(MU_NP / z_c, z_c, jet, alpha)
                   + _crit_sc2(theta, z_c, jet, alpha))
                + (theta < 2. * MU_NP)
                * (_crit_sc1(MU_NP / z_c, z_c, jet, alpha)
                   + _crit_sc2(2. * MU_NP, z_c, jet, ALPHA_FIXED)
                   + _crit_sc3(theta, z_c, jet, ALPHA_FIXED)))
        hard = ((theta > 2. * MU_NP) * _crit_hc1(theta, z_c, jet, alpha)
                + (theta < 2. * MU_NP)
                * (_crit_hc1(2. * MU_NP, z_c, jet, alpha)
                   + _crit_hc2(theta, z_c, jet, alpha)))
        return soft + hard


def radprime_crit_rc(theta, z_c, jet=QUARK, alpha=ALPHA_FIXED):
    """ref: running_coupling.py:326-347"""
    with np.errstate(all='ignore'):
        m = np.maximum(z_c, np.minimum(MU_NP / theta, 1. / 2.))
        sc = (np.log((1 + 2. * alpha * BETA_0 * np.log(theta / 2.))
                     / (1 + 2. * alpha * BETA_0 * np.log(m * theta)))
              / (2. * alpha * BETA_0)
              + np.log(m / z_c) / (1 + 2. * alpha * BETA_0 * np.log(MU_NP)))
        hc = 1. / (1. + 2. * alpha * BETA_0
                   * np.log(np.maximum(theta / 2., MU_NP)))
        cr2 = 2. * CF if jet == QUARK else 2. * CA
        casimir(jet)
        return alpha * (cr2 * sc + b_bar(z_c, jet) * hc) / (np.pi * theta)


def radprime_sub_rc(c, beta=2., jet=QUARK, max_radius=1.):
    """ref: running_coupling.py:350-380"""
    with np.errstate(all='ignore'):
        c = c / max_radius ** beta
        jac = 1. / max_radius ** beta
        alpha = alpha_one_loop(np.maximum(P_T * max_radius, 1.))
        m = np.maximum(c, np.minimum((MU_NP ** beta / c) ** (1. / (beta - 1.)), 1. / 2.))
        sc = (np.log((1 + 2. * alpha * BETA_0 * np.log(c / 2. ** (beta - 1.)) / beta)
                     / (1 + 2. * alpha * BETA_0 * np.log(c * m ** (beta - 1.)) / beta))
              * beta / ((beta - 1.) * 2. * alpha * BETA_0)
              + np.log(m / c) / (1 + 2. * alpha * BETA_0 * np.log(MU_NP)))
        hc = 1. / (1. + 2. * alpha * BETA_0 * np.log(
            np.maximum(2 ** (-(beta - 1.) / beta) * c ** (1 / beta), MU_NP)))
        cr2 = 2. * CF if jet == QUARK else 2. * CA
        casimir(jet)
        return jac * alpha * (cr2 * sc + b_bar(c, jet) * hc) / (np.pi * beta * c)


def pdf_crit_rc(z, theta, z_c, jet=QUARK):
    """ref: running_coupling.py:386-404"""
    with np.errstate(all='ignore'):
        return (splitting(z, jet, MLL) * alpha_running(z, theta)
                * np.exp(-1. * rad_crit_rc(theta, z_c, jet))
                / (np.pi * theta))


def pdf_sub_rc(c, beta=2., jet=QUARK, max_radius=1.):
    """ref: running_coupling.py:409-422"""
    with np.errstate(all='ignore'):
        return (radprime_sub_rc(c, beta, jet, max_radius=max_radius)
                * np.exp(-rad_sub_rc(c, beta, jet, max_radius=max_radius)))


# --------------------------------------------------------------------------
# Phase-space weights                             ref: numerics/weights.py
# --------------------------------------------------------------------------
def weight_radiator(z, theta, jet, fixed_coupling=True, acc=LL):
    """ref: numerics/weights.py:15-26"""
    alpha = ALPHA_FIXED if fixed_coupling else alpha_running(z, theta)
    return splitting(z, jet, acc) * alpha / (theta * np.pi)


def weight_critical(z, theta, z_c, jet=QUARK, fixed_coupling=True):
    """ref: numerics/weights.py:31-41"""
    if fixed_coupling:
        return pdf_crit_fc(z, theta, z_c, jet)
    return pdf_crit_rc(z, theta, z_c, jet)


def weight_precritical(z_pre, theta_crit, z_c, jet=QUARK, fixed_coupling=True):
    """ref: numerics/weights.py:43-58"""
    if not fixed_coupling:
        raise TypeError("Can only accept fixed coupling for now!")
    cr = casimir(jet)
    with np.errstate(all='ignore'):
        return (2. * cr * ALPHA_FIXED / np.pi
                * np.log(R0 / theta_crit) / z_pre
                * np.exp(-2. * cr * ALPHA_FIXED / np.pi
                         * np.log(R0 / theta_crit)
                         * np.log(z_c / (z_pre)))
                ) * (z_pre > 0.) * (z_pre < z_c)


def weight_precritical_zerobin(z_pre, theta_crit, z_c, jet, epsilon):
    """Zero-bin form used by the reference's all-emission comparison.
    ref: /root/reference/examples/tests/comparison_tests/
    comparison_utils_probabilistic.py:82-98"""
    assert 0 < epsilon < 1, "Weight requires valid cutoff for zero bin."
    cr = casimir(jet)
    with np.errstate(all='ignore'):
        zero = np.exp(-2. * cr * ALPHA_FIXED / np.pi
                      * np.log(R0 / theta_crit)
                      * np.log(z_c / epsilon)) * (z_pre == 0)
        w = weight_precritical(z_pre, theta_crit, z_c, jet, True) * (z_pre > 0)
    return zero + np.nan_to_num(w)


def weight_two_emission(crit, sub, z_c, beta, jet=QUARK, fixed_coupling=True):
    """Critical x subsequent weight; keeps ``csub = z_s th_s^b th_s^b`` verbatim.
    ref: numerics/weights.py:63-99"""
    if len(crit) != len(sub):
        raise AssertionError(
            "Must have the same amount of critical and subsequent samples!")
    z_c_em, th_c = crit[:, 0], crit[:, 1]
    z_s, th_s = sub[:, 0], sub[:, 1]
    csub = z_s * th_s ** beta * th_s ** beta
    if fixed_coupling:
        pc = pdf_crit_fc(z_c_em, th_c, z_c, jet).astype('float')
        ps = pdf_sub_fc(csub, beta, jet, max_radius=th_c)
    else:
        pc = pdf_crit_rc(z_c_em, th_c, z_c, jet).astype('float')
        ps = pdf_sub_rc(csub, beta, jet, max_radius=th_c)
    with np.errstate(all='ignore'):
        return pc * ps * (th_s ** beta * th_c ** beta)


# --------------------------------------------------------------------------
# Uniform random numbers
# --------------------------------------------------------------------------
def mt_uniforms(seed, n):
    """The first ``n`` values ``random.random()`` returns after
    ``random.seed(seed)`` -- vectorised by transplanting CPython's MT19937 state
    into NumPy's identical generator (both build doubles as
    (a>>5, b>>6) -> (a*2^26+b)/2^53).  ref: utils/montecarlo_utils.py:28,53"""
    gen = _pyrandom.Random(seed)
    version, state, _ = gen.getstate()
    rs = np.random.RandomState()
    rs.set_state(('MT19937', np.array(state[:-1], dtype=np.uint32), state[-1]))
    return rs.random_sample(int(n))


_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al., SC'11) on uint32 arrays.  This is the
    generator the CUDA path uses; the KAT in tests pins both to the published
    Random123 vectors."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                          hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def philox_uniforms53(seed, first_index, n, n_uniform):
    """FP64 uniforms of the CUDA path's ``f64`` mode: sample ``i`` draws block
    ``b`` = philox(counter=(i_lo, i_hi, b, 0), key=(seed_lo, seed_hi)) and
    uniform ``2b+j`` = ((w[2j]>>5)*2^26 + (w[2j+1]>>6)) / 2^53.
    Returns an (n, n_uniform) array.  (Product-side definition:
    jetmontecarlo_b200/csrc/jmc_philox.cuh.)"""
    idx = np.arange(n, dtype=np.uint64) + np.uint64(first_index)
    lo = idx & _MASK32
    hi = idx >> np.uint64(32)
    out = np.empty((n, n_uniform), dtype=np.float64)
    for b in range((n_uniform + 1) // 2):
        w = philox4x32_10(lo, hi, np.uint64(b), np.uint64(0),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(2):
            col = 2 * b + j
            if col < n_uniform:
                a = (w[2 * j] >> np.uint32(5)).astype(np.float64)
                c = (w[2 * j + 1] >> np.uint32(6)).astype(np.float64)
                out[:, col] = (a * 67108864.0 + c) / 9007199254740992.0
    return out


def philox_words(seed, first_index, n, block=0):
    """The four raw 32-bit words of block ``block`` for samples
    first_index .. first_index+n-1 (the ``f32`` mode consumes these directly)."""
    idx = np.arange(n, dtype=np.uint64) + np.uint64(first_index)
    return philox4x32_10(idx & _MASK32, idx >> np.uint64(32), np.uint64(block),
                         np.uint64(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


# --------------------------------------------------------------------------
# Samplers                                        ref: utils/montecarlo_utils.py:12-55,
#                                                      montecarlo/sampler.py:30-34,114-140
#                                                      numerics/radiators/samplers.py
# --------------------------------------------------------------------------
def _lin(lo, hi, u):
    """ref: utils/montecarlo_utils.py:28"""
    return lo + u * (hi - lo)


def _log(lo, hi, eps, u):
    """ref: utils/montecarlo_utils.py:53-55"""
    return lo + np.exp(np.log(eps) * u + np.log(hi - lo))


def check_sampler_args(method, epsilon, zc=None, radius=None, needs_zc=False):
    """Validity checks with the reference's messages.
    ref: montecarlo/sampler.py:65-81, numerics/radiators/samplers.py:12-18,116,298"""
    if needs_zc:
        z = 0 if zc is None else zc
        if not 0 < z < 1. / 2.:
            raise AssertionError("zc={zcut} is an invalid grooming parameter."
                                 .format(zcut=z))
    if radius is not None and not radius > 0:
        raise AssertionError("Jet radius must be greater than zero!")
    if method not in ('lin', 'log'):
        raise AssertionError(str(method) + "is not a supported sample method.")
    eps = 0 if epsilon is None else epsilon
    if method == 'log' and not 0 < eps < 1.:
        raise AssertionError("epsilon={epsilon} is invalid for {type} sampling."
                             .format(epsilon=eps, type=method))


def sample_simple(method, u, bounds=(0, 1), epsilon=0.):
    """simpleSampler.  u: (N,) uniforms.  Returns samples (N,), jac (N,), area.
    ref: montecarlo/sampler.py:114-140"""
    check_sampler_args(method, epsilon)
    if method == 'lin':
        x = _lin(bounds[0], bounds[1], u)
        return x, np.ones_like(x), bounds[1] - bounds[0]
    x = _log(bounds[0], bounds[1], epsilon, u)
    return x, x.copy(), np.log(1. / epsilon)


def sample_critical(method, u, zc, epsilon=0., radius=1.):
    """criticalSampler.  u: (N,2) uniforms, column 0 -> z, column 1 -> theta
    (the reference draws z first).  ref: numerics/radiators/samplers.py:30-62"""
    check_sampler_args(method, epsilon, zc, radius, needs_zc=True)
    if method == 'lin':
        z, th = _lin(zc, 1. / 2., u[:, 0]), _lin(0, radius, u[:, 1])
        return np.stack([z, th], 1), np.ones(len(z)), (1. / 2. - zc) * radius
    z, th = _log(zc, 1. / 2., epsilon, u[:, 0]), _log(0, radius, epsilon, u[:, 1])
    return np.stack([z, th], 1), (z - zc) * th, np.log(1. / epsilon) ** 2.


def sample_precritical(method, u, zc, epsilon=0.):
    """precriticalSampler.  ref: numerics/radiators/samplers.py:132-158"""
    check_sampler_args(method, epsilon, zc, needs_zc=True)
    if method == 'lin':
        z = _lin(0., zc, u)
        return z, np.ones_like(z), zc
    z = _log(0., zc, epsilon, u)
    return z, z.copy(), np.log(1. / epsilon)


def sample_ungroomed(method, u, epsilon=0., radius=1.):
    """ungroomedSampler.  ref: numerics/radiators/samplers.py:196-226"""
    check_sampler_args(method, epsilon, radius=radius)
    if method == 'lin':
        z, th = _lin(0., 1. / 2., u[:, 0]), _lin(0, radius, u[:, 1])
        return np.stack([z, th], 1), np.ones(len(z)), radius / 2.
    z, th = _log(0., 1. / 2., epsilon, u[:, 0]), _log(0, radius, epsilon, u[:, 1])
    return np.stack([z, th], 1), z * th, np.log(1. / epsilon) ** 2.


# --------------------------------------------------------------------------
# The reference's sampling LOOP, restated: one Python iteration per sample,
# one random.random() and scalar np.log / np.exp per coordinate, lists
# converted to arrays at the end.  This is what the reference actually runs
# (and where 98-99 % of its time goes); bench.py's CPU baseline times it.
#                                 ref: montecarlo/sampler.py:30-34,
#                                      utils/montecarlo_utils.py:12-55,
#                                      numerics/radiators/samplers.py:38-62,204-226
# --------------------------------------------------------------------------
def _get_lin_sample(rng, lo, hi):
    return lo + rng.random() * (hi - lo)


def _get_log_sample(rng, lo, hi, eps):
    return lo + np.exp(np.log(eps) * rng.random() + np.log(hi - lo))


def sample_loop_critical(method, n, zc, epsilon=0., seed=0, radius=1.):
    check_sampler_args(method, epsilon, zc, radius, needs_zc=True)
    rng = _pyrandom.Random(seed)
    samples, jacs = [], []
    for _ in range(n):
        if method == 'lin':
            z, th = _get_lin_sample(rng, zc, 1. / 2.), _get_lin_sample(rng, 0, radius)
            jac = 1.
        else:
            z, th = _get_log_sample(rng, zc, 1. / 2., epsilon), _get_log_sample(rng, 0, radius, epsilon)
            jac = (z - zc) * th
        samples.append([z, th])
        jacs.append(jac)
    return np.array(samples), np.array(jacs)


def sample_loop_ungroomed(method, n, epsilon=0., seed=0, radius=1.):
    check_sampler_args(method, epsilon, radius=radius)
    rng = _pyrandom.Random(seed)
    samples, jacs = [], []
    for _ in range(n):
        if method == 'lin':
            z, th = _get_lin_sample(rng, 0, 1. / 2.), _get_lin_sample(rng, 0, radius)
            jac = 1.
        else:
            z, th = _get_log_sample(rng, 0, 1. / 2., epsilon), _get_log_sample(rng, 0, radius, epsilon)
            jac = z * th
        samples.append([z, th])
        jacs.append(jac)
    return np.array(samples), np.array(jacs)


# --------------------------------------------------------------------------
# Integrator                                      ref: montecarlo/integrator.py:37-202
# --------------------------------------------------------------------------
MIN_LOG_BIN = 1e-15


def make_edges(numbins, observables, spacing, min_log_bin=MIN_LOG_BIN):
    """Bin edges and mid-points; ``numbins`` is the number of EDGES.
    ref: montecarlo/integrator.py:37-57"""
    if spacing == 'lin':
        edges = np.linspace(min(0, np.min(observables)), np.max(observables), numbins)
        mids = (edges[1:] + edges[:-1]) / 2.
    elif spacing == 'log':
        lo = np.log10(max(np.min(observables), min_log_bin))
        hi = np.log10(np.max(observables))
        edges = np.logspace(lo, hi, numbins)
        mids = np.exp((np.log(edges[1:]) + np.log(edges[:-1])) / 2.)
    else:
        raise ValueError(spacing)
    return edges, mids


def bin_index(obs, edges):
    """Per-sample statement of ``np.histogram(obs, edges)``: half-open bins,
    last bin closed, -1 for dropped (out of range or NaN).
    ref: numpy.histogram as called at montecarlo/integrator.py:97-101"""
    obs = np.asarray(obs, dtype=np.float64)
    idx = np.searchsorted(edges, obs, side='right').astype(np.int64) - 1
    idx[obs == edges[-1]] = len(edges) - 2
    bad = (idx < 0) | (idx > len(edges) - 2) | np.isnan(obs)
    idx[bad] = -1
    return idx.astype(np.int32)


def histograms(obs, weights, edges):
    """(sum w, sum w^2, count) per bin via NumPy itself.
    ref: montecarlo/integrator.py:97-101"""
    h, _ = np.histogram(obs, edges, weights=weights)
    h2, _ = np.histogram(obs, edges, weights=np.square(weights))
    n, _ = np.histogram(obs, edges)
    return h, h2, n


def density_from_hist(h, h2, n, edges, area, n_total):
    """ref: montecarlo/integrator.py:113-124"""
    widths = edges[1:] - edges[:-1]
    with np.errstate(all='ignore'):
        dens = (h / widths) * (area / n_total)
        err = (np.sqrt(h2) / widths) * (area / n_total)
    dens = np.where(n == 0, 0, dens)
    err = np.where(n == 0, 0, err)
    return dens, err


def cumulative(dens, err, edges, first_bc=None, last_bc=None):
    """Integral (len = numbins) and its error (len = numbins-1).
    ref: montecarlo/integrator.py:129-202"""
    use_last = last_bc is not None and first_bc is None
    use_first = first_bc is not None and last_bc is None
    if not (use_last or use_first):
        raise AssertionError("Integration requires unambiguous boundary conditions")
    widths = edges[1:] - edges[:-1]
    if use_first:
        cum = np.cumsum(dens * widths)
        return (np.array([first_bc, *(first_bc + cum)]), np.cumsum(err * widths))
    value, sign = last_bc
    if sign not in ('plus', 'minus'):
        raise AssertionError("Integration requires lastBinBndCond (plus or minus)")
    rev = np.cumsum((dens * widths)[::-1])[::-1]
    integral = [*(value + rev), value] if sign == 'plus' else [*(value - rev), value]
    return np.array(integral), np.cumsum((err * widths)[::-1])[::-1]


def integrate_hist(obs, weights, edges, area, first_bc=None, last_bc=None):
    """setDensity + integrate in one call; returns a dict."""
    h, h2, n = histograms(obs, weights, edges)
    dens, err = density_from_hist(h, h2, n, edges, area, len(obs))
    integ, ierr = cumulative(dens, err, edges, first_bc, last_bc)
    return dict(sum_w=h, sum_w2=h2, count=n, density=dens, density_err=err,
                integral=integ, integral_err=ierr)


def make_edges_2d(numbins, obs, spacing, min_log_bin=MIN_LOG_BIN):
    """ref: montecarlo/integrator.py:486-500 ('lin' bins of both axes start at min(0, min over BOTH arrays))"""
    bins = []
    for i in range(2):
        if spacing == 'lin':
            bins.append(np.linspace(min(0, np.min(obs)), np.max(obs[i]), numbins))
        else:
            bins.append(np.logspace(np.log10(max(np.min(obs[i]), min_log_bin)), np.log10(np.max(obs[i])), numbins))
    return bins


def integrate_hist_2d(x, y, weights, bins, area, first_bc=None, last_bc=None):
    """integrator_2d.setDensity + integrate + set_extended_integral.
    ref: montecarlo/integrator.py:463-467 (multidim_cumsum), :525-685"""
    h, _, _ = np.histogram2d(x, y, bins, weights=weights)
    h2, _, _ = np.histogram2d(x, y, bins, weights=np.square(weights))
    n, _, _ = np.histogram2d(x, y, bins)
    w1, w2 = np.meshgrid(bins[0][1:] - bins[0][:-1], bins[1][1:] - bins[1][:-1])
    areas = w1 * w2
    factor = area / (len(x) * areas.T)
    dens = np.where(n == 0, 0, h * factor).T
    err = np.where(n == 0, 0, np.sqrt(h2) * factor).T

    def mcumsum(a):
        out = a.cumsum(-1)
        np.cumsum(out, axis=-2, out=out)
        return out
    nx = len(bins[0])
    if first_bc is not None:
        integ = first_bc + mcumsum(dens * areas)
        ierr = mcumsum(err * areas)
        ext = np.array([np.ones(nx) * first_bc] + [np.append(first_bc, r) for r in integ]).T
        ext_err = np.array([np.zeros(nx)] + [np.append(0, r) for r in ierr]).T
    else:
        value, sign = last_bc
        rev = np.flip(mcumsum(np.flip(dens * areas)))
        integ = value + rev if sign == 'plus' else value - rev
        ierr = np.flip(mcumsum(np.flip(err * areas)))
        ext = np.array([np.append(r, value) for r in integ] + [np.ones(nx) * value]).T
        ext_err = np.array([np.append(0, r) for r in ierr] + [np.zeros(nx)]).T
    return dict(density=dens, density_err=err, integral=integ, integral_err=ierr, ext=ext, ext_err=ext_err,
                sum_w=h, sum_w2=h2, count=n)


def lin_log_mixed_list(lower, upper, num):
    """ref: utils/interpolation.py:18-41"""
    vals = np.logspace(np.log10(lower), np.log10(upper), int(num / 2) + 2)
    vals = np.append(vals, np.linspace(lower, upper, int(num / 2)))
    vals = np.sort(vals[1:-1])
    return vals[~np.isnan(vals)]


def vals_to_pdf(vals, num_bins, weights=None, bin_space='log', log_cutoff=None):
    """Weighted pdf on fixed bins; log bins get an underflow ("zero") bin.
    ref: utils/hist_utils.py:94-144"""
    vals = np.asarray(vals, dtype=float)
    weights = np.ones_like(vals) if weights is None else np.asarray(weights, dtype=float)
    good = np.isfinite(vals) & np.isfinite(weights)
    vals, weights = vals[good], weights[good]
    if bin_space == 'lin':
        bins = np.linspace(0, 1, num_bins)
        xs = (bins[:-1] + bins[1:]) / 2
    else:
        bins = np.logspace(np.log10(log_cutoff), 0, num_bins)
        bins = np.insert(bins, 0, 1e-100)
        xs = np.insert(np.sqrt(bins[1:-1] * bins[2:]), 0, 1e-50)
    pdf, _ = np.histogram(vals, bins, weights=weights)
    pdf = pdf / np.array(np.diff(bins), float) / pdf.sum()
    if bin_space == 'log':
        pdf = xs * pdf * np.log(10)
    return xs, pdf


# --------------------------------------------------------------------------
# Angularity shower with the MLL veto algorithm
#                                                 ref: utils/partonshower_utils.py:130-328
# --------------------------------------------------------------------------
ALPHA_VETO = 1.5


def shower_emissions(u_stream, beta, jet, acc, radius=1., cutoff=MU_NP, max_emissions=64,
                     azimuth_draw=True):
    """One jet: the ordered (z, theta) chain produced by gen_jets with
    split_soft=False.  ``u_stream`` is a callable returning the next uniform.
    With split_soft=False the 3-vector geometry is redundant: the opening angle
    of the daughters is theta and the softer momentum fraction is z.
    ``azimuth_draw``: the reference's Parton.split rotates the daughters about
    a random perpendicular axis and so burns one uniform per accepted emission
    (utils/vector_utils.py:127); it never influences (z, theta) but must be
    consumed to replay the reference's MT19937 stream.
    ref: utils/partonshower_utils.py:130-219 (split), :224-285 (recursion),
    :294-328 (driver)."""
    cr = casimir(jet)
    alpha = ALPHA_FIXED if acc == LL else ALPHA_VETO
    ang, z, theta = radius ** beta / 2., 1. / 2, radius
    chain = []
    while z * theta > cutoff and len(chain) < max_emissions:
        while True:
            r1, r2 = u_stream(), u_stream()
            ang_f = np.exp(-np.sqrt(np.log(2. * ang) ** 2.
                                    - np.pi * beta / (cr * alpha) * np.log(r1))) / 2.
            z = (2. * ang_f) ** r2 / 2.
            theta = (2. * ang_f) ** ((1 - r2) / beta)
            if acc == LL:
                break
            cut = (alpha_running(z, theta) * splitting(z, jet, MLL)
                   / (2. * cr * alpha / z))
            if cut > 1:
                raise ValueError("The pdf must be everywhere less than the"
                                 "proposed pdf!")
            if u_stream() < cut:
                break
            ang = ang_f
        ang = ang_f
        chain.append((z, theta))
        if azimuth_draw:
            u_stream()
    return chain


def ecf_from_chain_ungroomed(chain, beta, obs_acc=LL, n_emissions=1):
    """ref: utils/partonshower_utils.py:445-499"""
    cs = [1e-50] + [float(ecf_ungroomed(z, th, beta, obs_acc)) for z, th in chain]
    if n_emissions == 'all':
        return sum(cs)
    cs.sort(reverse=True)
    return sum(cs[:int(n_emissions)])


def ecf_from_chain_softdrop(chain, z_cut, beta, beta_sd=0., obs_acc=LL, n_emissions=1):
    """First emission passing z > z_cut theta^beta_sd is critical, later ones
    subsequent.  Per-jet statement of ref: utils/partonshower_utils.py:676-790
    with reproduce_approximations=True.  (The reference decrements its
    ``n_emissions`` argument in place at :765, so only its *first* jet uses
    two emissions; this function implements the per-jet rule the docstring
    states and tests pin both behaviours.)"""
    c_list, dropping = [], True
    for z, th in chain:
        if not dropping:
            c_list.append(float(ecf_ungroomed(z, th, beta, obs_acc)))
        elif z > z_cut * th ** beta_sd:
            dropping = False
            c_list.append(float(ecf_ungroomed(z, th, beta, obs_acc)))
    c_list += [1e-50, 1e-50]
    if n_emissions == 'all':
        return sum(c_list)
    if n_emissions == 1:
        return c_list[0]
    rest = sorted(c_list[1:], reverse=True)
    return sum([c_list[0]] + rest[:int(n_emissions) - 1])


def ecf_from_chain_rss(chain, z_cut, beta, f=1., obs_acc=LL, emission_type='precritsub', n_emissions=1,
                       few_pres=True):
    """Recursive-safe-subtraction grooming of one jet's (z, theta) chain: emissions softer than the remaining
    z_cut are pre-critical (they eat into z_cut and, for f != 1, rescale the hard branch), the first one that
    survives is critical, later ones are subsequent.  Returns None for a jet the reference drops (ecf < 0).
    ``emission_type`` is matched by substring exactly as the reference does ('precrit' also contains 'crit').
    ref: utils/partonshower_utils.py:501-674"""
    if z_cut == 0:
        return ecf_from_chain_ungroomed(chain, beta, obs_acc, n_emissions)
    assert f > 1 / 3, "f = " + str(f) + " is less than 1/3"

    def ecf(z, theta, zc):
        if obs_acc == LL:
            return (z - f * zc) * theta ** beta * (1. - (1. - f) * zc)
        return (z - f * zc) * theta ** beta * (1. - z - (1. - f) * zc)

    this_zcut, use_precrit = z_cut, True
    c_crit, z_pres, c_subs = 1e-50, [1e-50], [1e-50]
    for z, theta in chain:
        # scalar types of the reference: z is a ratio of math.sqrt magnitudes (Python float), theta comes out of
        # np.arccos (NumPy scalar, so theta**beta is NumPy's power, which squares exactly) -- bit-for-bit parity
        # with the golden vectors depends on it in about one jet per hundred
        z, theta = float(z), np.float64(theta)
        if few_pres:
            cutoff_zcut = this_zcut - max(z_pres)
            hard_scale = 1 - (1 - f) * max(z_pres) / f
        else:
            cutoff_zcut = this_zcut - sum(z_pres)
            hard_scale = np.prod([1 - (1 - f) * zp / f for zp in z_pres])
        valid_precrit = (z < this_zcut) if few_pres else (z < cutoff_zcut)
        if 'precrit' in emission_type and use_precrit and valid_precrit:
            z_pres.append(z)
        elif 'crit' in emission_type and z >= cutoff_zcut > 0.:
            c_crit = ecf(z, theta, cutoff_zcut) * hard_scale ** 2.
            this_zcut, use_precrit = 0., False
        elif 'sub' in emission_type and this_zcut == 0.:
            c_subs.append(ecf(z, theta, 0.) * hard_scale ** 2.)
    c_list = c_subs + [c_crit]
    if n_emissions == 'all':
        total = sum(c_list)
    else:
        c_list.sort(reverse=True)
        total = sum(c_list[:int(n_emissions)])
    return total if total >= 0 else None


# ---------------------------------------------------------------------------
# Inverse-transform sampling from a cdf given as a function
# ---------------------------------------------------------------------------
def inverse_transform(cdf_vals, x_vals, r):
    """x(r) by linear interpolation of the table (cdf -> x), extrapolating beyond it.
    ref: utils/montecarlo_utils.py:377-399 (scipy interp1d, which sorts its abscissa, here the cdf values)"""
    from scipy import interpolate
    return interpolate.interp1d(cdf_vals, x_vals, fill_value='extrapolate')(r)


def cdf_probe_points(domain):
    """The 1000-odd points on which the reference inspects a cdf.  ref: utils/montecarlo_utils.py:147-154"""
    if domain[0] < 0:
        return np.linspace(domain[0], domain[1], 1000)
    if domain[0] == 0:
        return np.array([0, *lin_log_mixed_list(domain[0] + 1e-20, domain[1], 1000)])
    return lin_log_mixed_list(domain[0], domain[1], 1000)


def samples_from_cdf_tabulated(cdf, num_samples, domain, rand, catch_turnpoint=False, force_monotone=False):
    """What samples_from_cdf does with a cdf that pynverse's ``inversefunc`` refuses (ValueError) -- the only part
    of it that can be executed here, pynverse (0.1.4.4) being absent: tabulate the cdf on the probe points, then
    in this order: monotone table -> inverse-transform samples of the table; cdf starting at 1 -> zero bin;
    negligible (< 1e-20) up to a point and monotone after it -> retry on the rest of the domain; cdf reaching 1 and
    turning back (``catch_turnpoint``) -> retry below the turning point; ``force_monotone`` -> drop the table
    entries after which the cdf decreases; otherwise ValueError.  ``rand(n)`` supplies the uniforms
    (np.random.rand in the reference).  Returns (samples, weights).
    ref: utils/montecarlo_utils.py:93-375 (the ``except ValueError`` branch, :142-318)"""
    pnts = cdf_probe_points(domain)
    cdf_vals = np.nan_to_num(cdf(pnts))
    monotone = cdf_vals[:-1] <= cdf_vals[1:]
    if monotone.all():
        samples = inverse_transform(cdf_vals, pnts, rand(num_samples))
        return samples, np.ones_like(samples)
    bad_low = pnts[:-1][~monotone]
    bad_cdf_low = cdf(np.unique(bad_low))
    if all(cdf_vals == 1.):
        return np.zeros(num_samples), np.ones(num_samples)
    if bad_low[0] == pnts[0] and bad_cdf_low[0] == 1:
        return np.zeros(num_samples), np.ones(num_samples)
    threshold = 1e-20
    if all(cdf_vals < threshold):
        samples = np.full(num_samples, domain[1])
        return samples, np.ones_like(samples)
    for i, (value, point) in enumerate(zip(cdf_vals, pnts)):
        if value > threshold:
            if not all(monotone[i:]):
                break
            return samples_from_cdf_tabulated(cdf, num_samples, [point, domain[1]], rand, catch_turnpoint,
                                              force_monotone)
    if bad_cdf_low[0] == 1 and catch_turnpoint:
        return samples_from_cdf_tabulated(cdf, num_samples, [domain[0], bad_low[0]], rand, catch_turnpoint,
                                          force_monotone)
    if force_monotone:
        samples = inverse_transform(cdf_vals[:-1][monotone], pnts[:-1][monotone], rand(num_samples))
        return samples, np.ones_like(samples)
    raise ValueError("cdf is not monotone on %s and neither catch_turnpoint nor force_monotone applies" % (list(domain),))


# ---------------------------------------------------------------------------
# Generic MonteCarloProcess (hit-or-miss box sampling) and the e+e- -> hadrons (LO) client with its EWOCs
# ---------------------------------------------------------------------------
def alpha_em_one_loop(mu):
    """ref: analytics/qcd_utils.py:108-112"""
    alpha_em_zmass = 1. / 127
    beta0_em = 1. / (3 * np.pi)
    return alpha_em_zmass / (1 + 2 * alpha_em_zmass * beta0_em * np.log(mu / M_Z))


def sum_square_quark_charges(energy):
    """ref: analytics/qcd_utils.py:72-96 (u, d above 0.01 GeV; s above 0.05; c above 2.5; b above 10; no top)"""
    total = 0
    if energy > 0.01:
        total += (1 / 3) ** 2
        total += (2 / 3) ** 2
    if energy > 0.05:
        total += (-1 / 3) ** 2
    if energy > 2.5:
        total += (2 / 3) ** 2
    if energy > 10.0:
        total += (-1 / 3) ** 2
    return total


def ee2hadrons_lo_prefactor(energy):
    """ref: analytics/ewocs/eec.py:9-19"""
    return 8 * energy ** 2. * alpha_one_loop(energy) * alpha_em_one_loop(energy) ** 2. * sum_square_quark_charges(energy)


def ee2h_in_phasespace(a, b):
    """0 < x_i < 1 for x_1 = 1 - a, x_2 = 1 - b, x_3 = 2 - x_1 - x_2.
    ref: examples/processes/ee2hadrons_LO.py:17-21"""
    x_1, x_2 = 1 - a, 1 - b
    x_3 = 2 - x_1 - x_2
    return (0 < x_1) & (x_1 < 1) & (0 < x_2) & (x_2 < 1) & (0 < x_3) & (x_3 < 1)


def ee2h_differential_xsec(a, b, prefactor):
    """The reference's expression VERBATIM: its second variable is the kinematic variable 1 - x_2 itself (the
    "1 -" is missing at examples/processes/ee2hadrons_LO.py:28), a deterministic quirk that is reproduced."""
    x_1, x_2 = 1 - a, b
    return prefactor * (x_1 ** 2. + x_2 ** 2.) / ((1. - x_1) * (1. - x_2))


def process_hit_or_miss(uniforms, num_samples, method, epsilon, bounds, in_phasespace, xsec):
    """MonteCarloProcess.addSample in a loop: draw every kinematic variable in its box, keep the point if it lies
    in the phase space (weight = jacobian * differential cross section), else count a rejection and try again.
    ``uniforms``: the sequence random.random() would produce.  Returns samples [N, d], weights [N], rejected.
    ref: montecarlo/process.py:117-143"""
    samples, weights, rejected, k = [], [], 0, 0
    log_eps = np.log(epsilon) if method == 'log' else None
    while len(samples) < num_samples:
        point, jac = [], 1.
        for lo, hi in bounds:
            u = uniforms[k]
            k += 1
            if method == 'lin':
                point.append(lo + u * (hi - lo))
            else:
                point.append(lo + np.exp(log_eps * u + np.log(hi - lo)))
                jac *= point[-1]
        if in_phasespace(*point):
            samples.append(point)
            weights.append(jac * xsec(*point))
        else:
            rejected += 1
    return np.array(samples), np.array(weights), rejected


def process_area(bounds, num_samples, rejected):
    """ref: montecarlo/process.py:154-175"""
    max_area = 1.
    for lo, hi in bounds:
        max_area *= hi - lo
    return max_area * (num_samples / (num_samples + rejected))


def ee2h_pair_observable(samples, which, kappa=None):
    """The pairwise observable of the three parton pairs (1,2), (2,3), (3,1), concatenated pair by pair as
    ee2hLO_PairwiseObservable.update does.  ref: examples/processes/ee2hadrons_LO.py:75-150"""
    x_1, x_2 = 1 - samples[:, 0], 1 - samples[:, 1]
    x_3 = 2 - x_1 - x_2
    out = []
    with np.errstate(all='ignore'):
        for x_i, x_j in ((x_1, x_2), (x_2, x_3), (x_3, x_1)):
            cos = 2. / (x_i * x_j) - 2. / x_i - 2. / x_j + 1
            if which == 'costheta':
                out.append(cos)
            elif which == 'theta':
                out.append(np.arccos(cos))
            elif which == 'm2':
                out.append(-(1 - x_i - x_j))
            elif which == 'gecf':
                out.append(2 ** (kappa / 2 - 2) * x_i * x_j * (1 - cos) ** (kappa / 2))
            elif which == 'z_i':
                out.append(x_i / 2)
            elif which == 'z_j':
                out.append(x_j / 2)
            else:
                raise ValueError(which)
    return np.concatenate(out)


def ewoc_weights(z1, z2, weights, energy_weights=(1, 1)):
    """ref: montecarlo/ewoc.py:23-27"""
    w1, w2 = energy_weights
    return z1 ** w1 * z2 ** w2 * weights


def ewoc_bins(observables, numbins, binspacing, minval=None, maxval=None):
    """generate_ewoc's bins: integrator.setBins on the observables with the optional end points joined.
    ref: montecarlo/ewoc.py:41-50"""
    obs = np.asarray(observables)
    if minval is not None:
        obs = np.concatenate([[minval], obs])
    if maxval is not None:
        obs = np.concatenate([obs, [maxval]])
    return make_edges(numbins, obs, binspacing)[0]
