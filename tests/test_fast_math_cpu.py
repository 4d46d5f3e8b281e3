"""The algebra behind the FP32 kernel, checked on the CPU against the oracle in FP64:
the log-domain observable, the "reference weight is not NaN" predicate (is_active in csrc/jmc_fast.cu) and the
affine bin index.  These tests restate the kernel's FORMULAS in NumPy -- they pin the derivations, the kernels
themselves are checked on the GPU (tests/test_gpu_fast.py)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import jmc_oracle as O

LN2 = np.log(2.)


def _c3_draws(n, seed, zc, eps):
    """critical x subsequent samples carried the way the kernel carries them: base-2 logs of delta = z - z_c,
    theta_c, z_s, theta_s"""
    u = O.philox_uniforms53(seed, 0, n, 4)
    l2e = np.log2(eps)
    l2d = u[:, 0] * l2e + np.log2(.5 - zc)
    l2tc = u[:, 1] * l2e
    l2zs = u[:, 2] * l2e - 1.
    l2ts = u[:, 3] * l2e
    return l2d, l2tc, l2zs, l2ts


@pytest.mark.parametrize("zc,beta,z_pre,f,acc", [(.1, 2., None, 0., 'MLL'), (.1, 2., 0., 1., 'LL'), (.05, 3., .02, .5, 'MLL'),
                                                 (.2, 1.5, .2, 0., 'LL')])
def test_log_domain_observable(zc, beta, z_pre, f, acc):
    """log2 C_groomed = beta log2 theta + log2((delta K + c0 K)(b0 - delta)),  K = (1 - z_cut)^-2,
    c0 = z_cut - f (z_cut - z_pre),  b0 = 1 - z_cut - (1-f)(z_cut - z_pre)  [MLL; LL: 1 - (1-f)(z_cut - z_pre)]"""
    z_pre = zc if z_pre is None else z_pre
    l2d, l2tc, _, _ = _c3_draws(200_000, 3, zc, 1e-10)
    delta, theta = np.exp2(l2d), np.exp2(l2tc)
    z = zc + delta
    ref = O.ecf_groomed(z, theta, zc, beta, z_pre=z_pre, f=f, acc=acc)
    zcl = zc - z_pre
    K = 1. / (1. - zc) ** 2
    c0K = (zc - f * zcl) * K
    b = (1. - zc - (1. - f) * zcl) - delta if acc == 'MLL' else 1. - (1. - f) * zcl
    l2obs = beta * l2tc + np.log2((delta * K + c0K) * b)
    # the reference forms zmin - f zcl from z = z_c + delta, i.e. with the rounding of z; compare where that
    # rounding is harmless (delta not below 1e-9 z_c), and bound it everywhere
    ok = delta > 1e-9 * zc
    assert_allclose(np.exp2(l2obs[ok]), ref[ok], rtol=1e-6)
    rel = np.abs(np.exp2(l2obs) / ref - 1.)
    assert np.all(rel <= 4e-16 * zc / delta + 1e-12)       # what is left is the reference's own cancellation in z - z_c


@pytest.mark.parametrize("jet", ['quark', 'gluon'])
@pytest.mark.parametrize("zc,beta", [(.1, 2.), (.05, 3.), (.2, 1.5)])
def test_landau_predicate_equals_nan_mask_of_the_reference_weight(jet, zc, beta):
    """twoEmissionWeight (running coupling) is NaN exactly where the kernel's predicate says "inactive":
      theta_c > max(exp(-1/k)/z_c, exp(-1/ca) M_Z/(mu_NP P_T))      [1 + Lambda(z_c theta_c) > 0, 1 + Lambda(mu_NP, alpha(theta_c)) > 0]
      C < 1                                                           [log(2 - 2C) in b_bar(C)]
      log2 C > c2  or  log2 theta_c + log2 C > c3                     [1 + Lambda(C, alpha(theta_c)) > 0]
    with C = csub / theta_c^beta, k = 2 alpha_fixed beta0, ca = 2 * 0.118 * beta0."""
    n, eps = 400_000, 1e-10
    l2d, l2tc, l2zs, l2ts = _c3_draws(n, 11, zc, eps)
    crit = np.stack([zc + np.exp2(l2d), np.exp2(l2tc)], 1)
    sub = np.stack([np.exp2(l2zs), np.exp2(l2ts)], 1)
    with np.errstate(all='ignore'):
        w = O.weight_two_emission(crit, sub, zc, beta, jet, False)
    ref_active = ~np.isnan(w)
    beta0 = (33. - 2. * O.N_F) / (12. * np.pi)
    k, ca = 2. * O.ALPHA_FIXED * beta0, 2. * O.ALPHA_MZ * beta0
    l2tc_min = max(-1. / k - np.log(zc), -1. / ca + np.log(O.M_Z) - np.log(O.MU_NP) - np.log(O.P_T)) / LN2
    c2 = (np.log(O.M_Z) - 1. / ca) / LN2
    c3 = c2 - np.log2(O.P_T)
    lcl2 = l2zs + 2. * beta * l2ts - beta * l2tc                       # log2(csub / theta_c^beta)
    pred = ((l2tc + lcl2 > c3) | (lcl2 > c2)) & (l2tc > l2tc_min) & (lcl2 < 0.)
    # samples within rounding of a threshold may fall either way; everything else must agree
    margin = np.minimum.reduce([np.abs(l2tc - l2tc_min), np.abs(lcl2), np.abs(lcl2 - c2), np.abs(l2tc + lcl2 - c3)])
    clear = margin > 1e-9
    assert clear.mean() > .999999
    assert np.array_equal(pred[clear], ref_active[clear])
    assert 1e-3 < ref_active.mean() < .02         # a fraction of a percent of the samples carry weight (0.3 % at z_c = .1, beta = 2)
    # non-NaN weights are finite, and zero exactly where the reference masks them (C >= 1/2: subsequent pdf support)
    assert np.all(np.isfinite(w[ref_active]))


def test_critical_predicate():
    """criticalEmissionWeight (running coupling) is NaN exactly where 1 + Lambda(z_c theta) <= 0"""
    zc, eps, n = .1, 1e-10, 300_000
    l2d, l2tc, _, _ = _c3_draws(n, 5, zc, eps)
    with np.errstate(all='ignore'):
        w = O.weight_critical(zc + np.exp2(l2d), np.exp2(l2tc), zc, 'quark', False)
    beta0 = (33. - 2. * O.N_F) / (12. * np.pi)
    k = 2. * O.ALPHA_FIXED * beta0
    pred = 1. + k * (LN2 * l2tc + np.log(zc)) > 0.
    clear = np.abs(1. + k * (LN2 * l2tc + np.log(zc))) > 1e-9
    assert np.array_equal(pred[clear], ~np.isnan(w[clear]))
    assert .3 < pred.mean() < .4                                        # 35 % of the critical samples are above the Landau scale


def test_affine_log_bin_index_is_numpy_histogram():
    """np.logspace edges: bin = floor((log2 obs - log2 e0) / d), in range iff 0 <= that < nb -- identical to
    np.histogram except for observables within rounding of an edge"""
    rng = np.random.default_rng(2)
    edges = np.logspace(-11, np.log10(.5), 100)
    obs = np.exp(rng.uniform(np.log(1e-13), np.log(2.), 2_000_000))
    dom = np.log2(edges)
    d = (dom[-1] - dom[0]) / (len(edges) - 1)
    f = np.log2(obs) / d - dom[0] / d
    ok = (f >= 0) & (f < len(edges) - 1)
    idx = np.floor(f).astype(np.int64)
    ref = O.bin_index(obs, edges)                                       # -1 where np.histogram drops the sample
    near_edge = np.abs(f - np.round(f)) < 1e-9
    assert near_edge.mean() < 1e-6
    sel = ~near_edge
    assert np.array_equal(np.where(ok, idx, -1)[sel], ref[sel])
    # the single unsigned compare of the FP32 bit pattern used by the kernel is that same range test
    f32 = f.astype(np.float32)
    bits = f32.view(np.uint32)
    nb_bits = np.float32(len(edges) - 1).view(np.uint32)
    assert np.array_equal(bits < nb_bits, (f32 >= 0) & (f32 < np.float32(len(edges) - 1)) & ~(np.signbit(f32)))
    for special in (np.nan, -np.inf, np.inf, -0.0, -1e-30):
        assert not (np.float32(special).view(np.uint32) < nb_bits)
