"""GPU tests of the FP32 log-space mode of the fused kernel (the throughput path).

Parity contract for fresh samples (BASELINE.json north_star): pdf and cdf agree
with the CPU output within 3 sigma per bin and the cdf within 1e-3 absolute at
1e7 samples.  On top of that statistical statement the per-sample dump
(jmc_dump) lets the oracle evaluate the FP64 reference formulas on exactly the
samples the FP32 path drew, so the active-branch / Landau-predicate logic is
checked sample by sample:
  * observable within 2e-5 relative (FP32 log2/exp2 round trip),
  * weight within 2e-4 relative where the reference weight is finite (plus 3e-5/|x| next to the kinematic
    end point x = ln(2C/theta_c^beta) -> 0 of the fixed-coupling pdfs, where the weight itself vanishes
    like |x| and FP32 rounding of the logarithms is an absolute error on x),
  * zero weight exactly where the reference is NaN (np.nan_to_num), except
    within 1e-5 (relative, in the log argument) of a Landau threshold.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu

EPS, ZC, BETA = 1e-10, .1, 2.


def _endpoint_x(kind, samples, fc, beta=BETA):
    """|x| of the fixed-coupling subsequent pdf (inf where there is no such end point)"""
    thc, zs, ths = samples[:, 1], samples[:, 2], samples[:, 3]
    with np.errstate(all='ignore'):
        if kind == 'crit_sub' and fc:
            return np.abs(np.log(2. * zs * ths ** (2 * beta) / thc ** beta))
        if kind == 'pre_crit_sub':
            return np.abs(np.log(2. * zs))
    return np.full(len(samples), np.inf)


def _oracle_on_dump(kind, samples, jet, fc, acc, zc=ZC, beta=BETA, eps=EPS, z_pre=None, f=None, theta_scale=1.):
    zc_, thc, zs, ths, zpre, zbu, delta = (samples[:, k] for k in range(7))
    with np.errstate(all='ignore'):
        if kind == 'radiator':
            th_em = ths            # the dump holds theta_em = theta * theta_scale
            obs = O.ecf_ungroomed(zs, th_em, beta, acc)
            w = O.weight_radiator(zs, th_em, jet, fc, acc) * (zs * th_em)
            return obs, w
        jc = delta * thc          # (z - z_c) theta with the sampled delta: no cancellation
        if kind == 'crit':
            obs = O.ecf_groomed(zc_, thc, zc, beta, z_pre=0. if z_pre is None else z_pre,
                                f=1. if f is None else f, acc=acc)
            if z_pre in (None, 0.) and f in (None, 1.):
                # zmin - z_c is delta itself; recompute without cancellation for the comparison
                obs = (delta * thc ** beta / (1. - zc) ** 2. if acc == 'LL'
                       else delta * (1. - zc_) * thc ** beta / (1. - zc) ** 2.)
            w = O.weight_critical(zc_, thc, zc, jet, fc) * jc
            return obs, w
        js = zs * ths
        if kind == 'crit_sub':
            csub = zs * ths ** beta * ths ** beta
            obs = np.maximum(O.ecf_groomed(zc_, thc, zc, beta, z_pre=zc, f=0., acc=acc), csub)
            crit, sub = np.stack([zc_, thc], 1), np.stack([zs, ths], 1)
            w = O.weight_two_emission(crit, sub, zc, beta, jet, fc) * jc * js
            return obs, w
        # pre_crit_sub: the dump's z_pre is already 0 for zero-binned samples
        zp = zpre
        c_sub = zs * thc ** beta
        a = delta + np.where(zp == 0, 0., zp)       # f = 1: zmin - (z_c - z_pre) = delta + z_pre
        obs = np.maximum(a * thc ** beta / (1. - zc) ** 2., c_sub)
        w = (O.weight_critical(zc_, thc, zc, jet, True) * O.pdf_sub_fc(c_sub / thc ** beta, beta, jet)
             * O.weight_precritical_zerobin(zp, thc, zc, jet, eps))
        jp = np.where(zp == 0, 1., zp)
        return obs, w * jc * js * jp


CASES = [
    ('radiator', 'quark', True, 'LL'), ('radiator', 'gluon', False, 'MLL'), ('radiator', 'quark', False, 'MLL'),
    ('crit', 'quark', True, 'LL'), ('crit', 'quark', False, 'LL'), ('crit', 'gluon', False, 'MLL'),
    ('crit_sub', 'quark', False, 'MLL'), ('crit_sub', 'gluon', False, 'MLL'), ('crit_sub', 'quark', True, 'MLL'),
    ('crit_sub', 'gluon', True, 'LL'),
    ('pre_crit_sub', 'quark', True, 'LL'), ('pre_crit_sub', 'gluon', True, 'LL'),
]


def _integrand(kind, jet, fc, acc, eps=EPS, **kw):
    from jetmontecarlo_b200.fused import Integrand
    if kind == 'radiator':
        return Integrand.radiator('log', eps, beta=BETA, jet_type=jet, fixedcoupling=fc, obs_acc=acc, split_acc=acc,
                                  nan_to_num=True, **kw)
    if kind == 'crit':
        return Integrand.critical('log', ZC, eps, beta=BETA, jet_type=jet, fixedcoupling=fc, obs_acc=acc,
                                  nan_to_num=True)
    if kind == 'crit_sub':
        return Integrand.crit_sub('log', ZC, eps, beta=BETA, jet_type=jet, fixedcoupling=fc, obs_acc=acc)
    return Integrand.pre_crit_sub('log', ZC, eps, beta=BETA, jet_type=jet, nan_to_num=True)


@pytest.mark.parametrize("kind,jet,fc,acc", CASES)
@pytest.mark.parametrize("eps", [1e-10, 1e-3])
def test_fast_per_sample_vs_oracle(cuda, kind, jet, fc, acc, eps):
    import gpu_helpers as H
    if kind == 'pre_crit_sub' and eps != 1e-10:
        eps = 1e-6
    ig = _integrand(kind, jet, fc, acc, eps)
    n = 400000
    samples, obs, w = H.dump(ig, n, 31337, 'f32')
    obs_ref, w_ref = _oracle_on_dump(kind, samples, jet, fc, acc, eps=eps)
    # observable
    ok = np.isfinite(obs_ref) & (obs_ref > 0)
    assert ok.mean() > .999
    assert_allclose(obs[ok], obs_ref[ok], rtol=2e-5)
    # weights: zero exactly where the reference is NaN ...
    nan_ref = np.isnan(w_ref)
    mism = (nan_ref != (w == 0)) & ~((w_ref == 0) & (w == 0))
    # ... except next to a Landau threshold (FP32 vs FP64 on the log argument) and FP32 underflow of tiny weights
    tiny = np.abs(np.nan_to_num(w_ref)) < 1e-30
    assert (mism & ~tiny).sum() <= max(3, 2e-5 * n), "%d samples disagree on NaN-ness" % (mism & ~tiny).sum()
    good = ~nan_ref & (w != 0) & (np.abs(w_ref) > 1e-30)
    if kind in ('crit', 'crit_sub') and not fc and eps < 1e-5:
        assert nan_ref.mean() > .3            # the Landau region is exercised
        assert good.sum() > 100
    # backward-error statement: the FP32 weight equals the reference function within 2e-4 relative, or is the
    # reference function at inputs moved by ~1e-5 relative (FP32 rounding of the carried logarithms).  The
    # second term only matters where the weight vanishes or changes sign (kinematic end point x -> 0 of the
    # fixed-coupling pdf, zero of R' of the running-coupling subsequent radiator near C = 1/2).
    moved = samples.copy()
    moved[:, 2] *= 1. + 1e-5
    _, w_moved = _oracle_on_dump(kind, moved, jet, fc, acc, eps=eps)
    tol = 2e-4 * np.abs(w_ref) + 3. * np.abs(np.nan_to_num(w_moved) - np.nan_to_num(w_ref)) \
        + 3e-5 / _endpoint_x(kind, samples, fc) * np.abs(w_ref)
    assert np.all(np.abs(w[good] - w_ref[good]) <= tol[good]), \
        "max rel weight error %.3g" % np.max(np.abs(w[good] / w_ref[good] - 1))
    # and in aggregate the weighted totals agree much better than per sample
    assert abs(w[good].sum() / w_ref[good].sum() - 1) < 2e-5


def test_fast_radiator_theta_scale(cuda):
    """gen_crit_sub_num_rad rescaling: theta_em = theta * theta_crit, jac * theta_crit"""
    import gpu_helpers as H
    ig = _integrand('radiator', 'gluon', False, 'MLL', theta_scale=.01)
    samples, obs, w = H.dump(ig, 100000, 5, 'f32')
    obs_ref, w_ref = _oracle_on_dump('radiator', samples, 'gluon', False, 'MLL')
    assert_allclose(obs, obs_ref, rtol=2e-5)
    assert_allclose(w, w_ref, rtol=2e-4)
    assert samples[:, 3].max() <= .01 * (1 + 1e-6)


@pytest.mark.parametrize("kind,jet,fc,acc", [CASES[0], CASES[1], CASES[4], CASES[6], CASES[8], CASES[10]])
def test_fast_histogram_equals_histogram_of_its_dump(cuda, kind, jet, fc, acc):
    """the fused FP32 kernel bins exactly the per-sample values it would dump: counts are identical to
    np.histogram of the dumped observable up to FP32 edge ties, sums agree to FP64 summation order"""
    import gpu_helpers as H
    ig = _integrand(kind, jet, fc, acc)
    n, seed = 300000, 77
    edges = np.logspace(-11, np.log10(.5), 100)
    _, obs, w = H.dump(ig, n, seed, 'f32')
    r = H.run(ig, n, seed, edges, 'f32')
    cnt = r['count'].cpu().numpy()
    # restate the kernel's binning exactly: np.logspace edges -> bin = floor of an affine map of log2(obs) in
    # FP32 (one FMA); in range = mapped value in [0, number of bins)
    l2 = np.log2(obs).astype(np.float32)
    dom = np.log2(edges)
    d = (dom[-1] - dom[0]) / (len(edges) - 1)
    # the host anchors the map at the FP32 first edge and lowers the slope ulp by ulp until the FP32 last edge maps
    # below the number of bins (make_fast_params)
    nb_f, e_lo, e_hi = np.float32(len(edges) - 1), np.float32(dom[0]), np.float32(dom[-1])
    fmaf = lambda x, a, c: np.float32(np.float64(x) * np.float64(a) + np.float64(c))
    inv_d = np.float32(1. / d)
    for _ in range(4096):
        mb0 = np.float32(-np.float64(e_lo) * np.float64(inv_d))
        while fmaf(e_lo, inv_d, mb0) < 0:
            mb0 = np.nextafter(mb0, np.float32(np.inf))
        if fmaf(e_hi, inv_d, mb0) < nb_f:
            break
        inv_d = np.nextafter(inv_d, np.float32(0))
    fma = (l2.astype(np.float64) * np.float64(inv_d) + np.float64(mb0)).astype(np.float32)
    keep = (fma >= 0) & (fma < len(edges) - 1)
    idx = np.floor(np.where(keep, fma, 0)).astype(np.int64)
    ref_cnt = np.bincount(idx[keep], minlength=len(edges) - 1)
    assert_array_equal(cnt, ref_cnt)
    assert abs(cnt - np.histogram(obs, edges)[0]).sum() <= 1e-4 * n     # and FP64 binning differs only at edge ties
    ref_sw = np.bincount(idx[keep], weights=w[keep], minlength=len(edges) - 1)
    ref_sw2 = np.bincount(idx[keep], weights=w[keep] ** 2, minlength=len(edges) - 1)
    # the dump and the histogram kernel are separate instantiations of the same FP32 code; FMA contraction
    # may differ between them in the last place of a weight, hence FP32-ulp level agreement of the sums
    assert_allclose(r['sum_w'].cpu().numpy(), ref_sw, rtol=2e-6, atol=1e-7 * np.abs(ref_sw).max())
    assert_allclose(r['sum_w2'].cpu().numpy(), ref_sw2, rtol=4e-6, atol=1e-7 * np.abs(ref_sw2).max())


def _pulls(a, ea, b, eb):
    err = np.sqrt(ea ** 2 + eb ** 2)
    m = err > 0
    return (a[m] - b[m]) / err[m]


@pytest.mark.parametrize("kind,jet,fc,acc", [CASES[0], CASES[1], CASES[3], CASES[4], CASES[6], CASES[8], CASES[10]])
def test_fast_vs_exact_fresh_samples(cuda, kind, jet, fc, acc):
    """fresh-sample agreement: FP32 path vs FP64 path on DIFFERENT streams, 1e7 samples each:
    <= 3 sigma per bin (allowing the expected handful of 3-sigma outliers over ~200 bins; none above 5),
    cdf within 1e-3 absolute"""
    from jetmontecarlo_b200 import fused
    ig = _integrand(kind, jet, fc, acc)
    n = 10_000_000
    edges = np.logspace(-11, np.log10(.5), 100)
    bc = [0., 'plus'] if kind == 'radiator' else [1., 'minus']
    a = fused.integrate(ig, n, edges=edges, seed=1, precision='f32', last_bc=bc)
    b = fused.integrate(ig, n, edges=edges, seed=2, precision='f64', last_bc=bc)
    p = _pulls(a.density, a.densityErr, b.density, b.densityErr)
    assert np.abs(p).max() < 5. and (np.abs(p) > 3.).sum() <= 3, "pdf pulls: max %.2f" % np.abs(p).max()
    ia, ib = np.asarray(a.integral), np.asarray(b.integral)
    q = _pulls(ia[:-1], a.integralErr, ib[:-1], b.integralErr)
    assert np.abs(q).max() < 4.
    if kind != 'radiator':                      # cdfs (the radiator integral is not bounded by 1)
        assert np.abs(ia - ib).max() < 1e-3 + 3 * np.sqrt(a.integralErr.max() ** 2 + b.integralErr.max() ** 2)


def test_fast_same_stream_is_much_closer_than_statistics(cuda):
    """same sample indices in both precisions (24-bit vs 53-bit uniforms from the same Philox words):
    this is NOT the same sample set, but the headline C3 cdf must agree within errors; and the FP32 run is
    reproducible bit for bit and additive over index ranges"""
    import gpu_helpers as H
    ig = _integrand('crit_sub', 'quark', False, 'MLL')
    edges = np.logspace(-11, np.log10(.5), 100)
    n = 4_000_000
    r1 = H.run(ig, n, 9, edges, 'f32')
    r2 = H.run(ig, n, 9, edges, 'f32')
    assert_array_equal(r1['count'].cpu().numpy(), r2['count'].cpu().numpy())
    # (sum w, sum w^2) of a block live in FP32 slots updated by compare-and-swap, in whatever order the warps arrive;
    # the block totals are summed in FP64: run-to-run differences are FP32 rounding of the few adds a slot sees
    assert_allclose(r1['sum_w'].cpu().numpy(), r2['sum_w'].cpu().numpy(), rtol=2e-6)
    acc = None
    for first, cnt in ((0, 1_500_000), (1_500_000, 2_500_000)):
        rr = H.run(ig, cnt, 9, edges, 'f32', first=first, accumulate_into=acc)
        acc = (rr['sum_w'], rr['sum_w2'], rr['count'])
    assert_array_equal(acc[2].cpu().numpy(), r1['count'].cpu().numpy())
    assert_allclose(acc[0].cpu().numpy(), r1['sum_w'].cpu().numpy(), rtol=2e-6)


@pytest.mark.parametrize("kind,jet,fc,acc", [CASES[8], CASES[6], CASES[0]])
def test_fast_index_range_across_2_to_the_32(cuda, kind, jet, fc, acc):
    """the kernel's sample index is 32-bit arithmetic inside a launch; the host splits a job where the global
    index crosses a multiple of 2^32.  A range straddling the boundary must equal the histogram of the dump of
    the same 64-bit indices, and the sum of its two halves."""
    import gpu_helpers as H
    ig = _integrand(kind, jet, fc, acc)
    edges = np.logspace(-11, np.log10(.5), 100)
    first, n, seed = 2 ** 32 - 70_001, 200_000, 5
    r = H.run(ig, n, seed, edges, 'f32', first=first)
    _, obs, w = H.dump(ig, n, seed, 'f32', first=first)
    cnt = r['count'].cpu().numpy()
    assert abs(cnt - np.histogram(obs, edges)[0]).sum() <= 1e-4 * n
    ref_sw = np.histogram(obs, edges, weights=w)[0]
    assert_allclose(r['sum_w'].cpu().numpy(), ref_sw, rtol=1e-3, atol=1e-4 * np.abs(ref_sw).max())
    acc_ = None
    for f0, c in ((first, 70_001), (2 ** 32, n - 70_001)):
        rr = H.run(ig, c, seed, edges, 'f32', first=f0, accumulate_into=acc_)
        acc_ = (rr['sum_w'], rr['sum_w2'], rr['count'])
    assert_array_equal(acc_[2].cpu().numpy(), cnt)
    assert_allclose(acc_[0].cpu().numpy(), r['sum_w'].cpu().numpy(), rtol=1e-6)     # a sample that moves between the main loop and the ragged-end trips may change by an FP32 ulp
    # and the samples on both sides of the boundary are different samples (the high word enters the counter)
    _, obs_lo, _ = H.dump(ig, 1000, seed, 'f32', first=0)
    _, obs_hi, _ = H.dump(ig, 1000, seed, 'f32', first=2 ** 32)
    assert not np.array_equal(obs_lo, obs_hi)


def test_fast_nonuniform_and_linear_edges(cuda):
    """general edge arrays: non-uniform (binary search), linear edges starting at 0 (linear compare domain),
    5000 edges (params.py NUM_RAD_BINS)"""
    import gpu_helpers as H
    ig = _integrand('crit', 'quark', True, 'LL')
    n, seed = 500000, 3
    _, obs, w = H.dump(ig, n, seed, 'f32')
    rng = np.random.RandomState(0)
    for edges in (np.sort(np.exp(rng.uniform(np.log(1e-12), np.log(.6), 40))),
                  np.linspace(0., .3, 50),
                  np.logspace(-12, 0, 5000)):
        r = H.run(ig, n, seed, edges, 'f32')
        cnt = r['count'].cpu().numpy()
        ref = np.histogram(obs, edges)[0]
        assert abs(cnt - ref).sum() <= 2e-4 * n + 2
        assert_allclose(r['sum_w'].cpu().numpy().sum(), np.histogram(obs, edges, weights=w)[0].sum(), rtol=1e-6)


def test_fast_minmax_and_unsupported(cuda):
    import gpu_helpers as H
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.fused import Integrand
    ig = _integrand('crit_sub', 'quark', True, 'MLL')
    n = 200000
    _, obs, _ = H.dump(ig, n, 4, 'f32')
    r = H.run(ig, n, 4, None, 'f32', want_minmax=True)
    assert_allclose(r['minmax'], [obs.min(), obs.max()], rtol=1e-6)
    with pytest.raises(_lib.JmcError, match="log-sampling path"):
        H.run(Integrand.critical('lin', .1), 10, 0, np.linspace(0, 1, 5), 'f32')


def test_c3_headline_against_cpu_oracle(cuda):
    """the headline workload end to end: FP32 fused path vs the oracle on CPU-generated (MT19937) samples"""
    from jetmontecarlo_b200 import fused
    n_cpu, n_gpu = 1_000_000, 20_000_000
    edges = np.logspace(np.log10(EPS) - 1, np.log10(.5), 100)
    for fc in (False, True):
        u = O.mt_uniforms(20260101 + int(fc), 4 * n_cpu)
        crit, jc, ac = O.sample_critical('log', u[:2 * n_cpu].reshape(n_cpu, 2), ZC, EPS)
        sub, js, as_ = O.sample_ungroomed('log', u[2 * n_cpu:].reshape(n_cpu, 2), EPS)
        with np.errstate(all='ignore'):
            csub = sub[:, 0] * sub[:, 1] ** BETA * sub[:, 1] ** BETA
            obs = np.maximum(O.ecf_groomed(crit[:, 0], crit[:, 1], ZC, BETA, z_pre=ZC, f=0., acc='MLL'), csub)
            w = np.nan_to_num(O.weight_two_emission(crit, sub, ZC, BETA, 'quark', fc)) * jc * js
        ref = O.integrate_hist(obs, w, edges, ac * as_, last_bc=[1., 'minus'])
        ig = _integrand('crit_sub', 'quark', fc, 'MLL')
        it = fused.integrate(ig, n_gpu, edges=edges, seed=5, precision='f32', last_bc=[1., 'minus'])
        p = _pulls(it.density, it.densityErr, ref['density'], ref['density_err'])
        assert np.abs(p).max() < 5. and (np.abs(p) > 3.).sum() <= 2
        q = _pulls(np.asarray(it.integral)[:-1], it.integralErr, ref['integral'][:-1], ref['integral_err'])
        assert np.abs(q).max() < 4.
        # SURVEY 8(d) caveat values: total integral 0.2705 (rc) / 0.5881 (fc) at eps = 1e-10
        total = 1. - it.integral[0]
        assert abs(total - (.5881 if fc else .2705)) < 4 * it.integralErr[0] + 2e-3


def test_full_size_properties(cuda):
    """BASELINE sizes (1e9 samples for C3, 1e10 for C4) through size-independent properties: every sample is either
    binned or out of range (counts + dropped = N with the dropped fraction known from a small dump), additivity
    over index ranges is exact in the counts, the cdf is monotone, ends at the boundary value, and agrees with a
    100x smaller job within errors"""
    import gpu_helpers as H
    from jetmontecarlo_b200 import fused
    edges = np.logspace(-11, np.log10(.5), 100)
    ig = _integrand('crit_sub', 'quark', False, 'MLL')
    n = 1_000_000_000
    r = H.run(ig, n, 5, edges, 'f32')
    cnt = r['count'].cpu().numpy()
    _, obs, _ = H.dump(ig, 2_000_000, 5, 'f32')
    in_range = ((obs >= edges[0]) & (obs <= edges[-1])).mean()
    assert abs(cnt.sum() / n - in_range) < 5 * np.sqrt(in_range * (1 - in_range) / 2e6)
    acc = None
    for first, c in ((0, 400_000_000), (400_000_000, 600_000_000)):
        rr = H.run(ig, c, 5, edges, 'f32', first=first, accumulate_into=acc)
        acc = (rr['sum_w'], rr['sum_w2'], rr['count'])
    assert_array_equal(acc[2].cpu().numpy(), cnt)
    # (the main loop and the ragged-end trips are separate instantiations of the FP32 code, so a sample that
    # changes position between them may change its weight in the last place: 1e-7 relative on a few samples)
    assert_allclose(acc[0].cpu().numpy(), r['sum_w'].cpu().numpy(), rtol=1e-6)
    big = fused.integrate(ig, n, edges=edges, seed=5, precision='f32', last_bc=[1., 'minus'])
    small = fused.integrate(ig, n // 100, edges=edges, seed=6, precision='f32', last_bc=[1., 'minus'])
    integ = np.asarray(big.integral)
    # (not monotone in its last bins: the reference's subRadPrimeAnalytic is negative for 1/2 < C < 1, where its
    # radiator is masked to 0, so the verbatim weight has a small negative tail -- reproduced, not "fixed")
    assert integ[-1] == 1. and 0 < integ[0] < 1 and integ.max() < 1.001, (integ[0], integ.max())
    assert np.all(np.diff(integ)[:-3] >= -1e-9), np.diff(integ).min()
    fcjob = fused.integrate(_integrand('crit_sub', 'quark', True, 'MLL'), n, edges=edges, seed=5, precision='f32',
                            last_bc=[1., 'minus'])
    assert np.all(np.diff(np.asarray(fcjob.integral)) >= 0) and fcjob.integral[-1] == 1.
    pull = (integ[:-1] - np.asarray(small.integral)[:-1]) / np.sqrt(big.integralErr ** 2 + small.integralErr ** 2)
    # where the small job still has statistics above the bin (its error estimate is meaningless on a handful of
    # heavy-tailed weights)
    enough = np.cumsum(small._count.cpu().numpy()[::-1])[::-1] >= 20000
    assert enough.sum() > 50 and np.abs(pull[enough]).max() < 4., np.abs(pull[enough]).max()
    # C4 at 1e10 samples (the size BASELINE quotes for 8 GPUs; 8 ms-scale chunks on one)
    ig4 = _integrand('pre_crit_sub', 'quark', True, 'LL')
    n4 = 10_000_000_000
    it = fused.integrate(ig4, n4, edges=np.logspace(-12, np.log10(.5), 100), seed=1, precision='f32',
                         last_bc=[1., 'minus'])
    assert .3 * n4 < int(it._count.sum().item()) <= n4
    integ = np.asarray(it.integral)
    assert integ[-1] == 1. and np.all(np.diff(integ) >= 0)


@pytest.mark.parametrize("kind,jet,fc,acc", [CASES[0], CASES[4], CASES[6], CASES[8], CASES[10],
                                             ('crit_sub_counts_on_request', 'quark', False, 'MLL')])
def test_fast_chunked_work_covers_every_sample_once(cuda, kind, jet, fc, acc):
    """Large jobs hand work out in chunks from a global counter (re-armed by the last block), small ones give every
    warp a fixed share: one job over an odd index range equals the sum of a large, a small and a tiny job over the
    same range, counts exactly -- every sample is drawn once whichever path took it, launch after launch."""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    ig = (_integrand(kind, jet, fc, acc) if kind != 'crit_sub_counts_on_request' else
          Integrand.crit_sub('log', ZC, EPS, beta=BETA, jet_type=jet, fixedcoupling=fc, obs_acc=acc, counts='weighted'))
    edges = np.logspace(-11, np.log10(.5), 100)
    first, n = 1, 33_333_337
    whole = H.run(ig, n, 9, edges, 'f32', first=first)
    again = H.run(ig, n, 9, edges, 'f32', first=first)
    assert_array_equal(whole['count'].cpu().numpy(), again['count'].cpu().numpy())
    acc_bufs, at = None, first
    for piece in (30_000_001, 3_333_000, 336):
        rr = H.run(ig, piece, 9, edges, 'f32', first=at, accumulate_into=acc_bufs)
        acc_bufs = (rr['sum_w'], rr['sum_w2'], rr['count'])
        at += piece
    assert at == first + n
    assert_array_equal(acc_bufs[2].cpu().numpy(), whole['count'].cpu().numpy())
    assert_allclose(acc_bufs[0].cpu().numpy(), whole['sum_w'].cpu().numpy(), rtol=1e-6, atol=1e-12)
    assert int(whole['count'].sum().item()) > 0


def test_batched_grid_equals_point_by_point(cuda):
    """jmc_run_batch: a theta_crit grid in one launch == one jmc_run per point (same streams: counts exact), with
    given edges and with the reference's data-dependent log bins; and gen_crit_sub_num_rad_fused reproduces the
    analytic fixed-coupling radiator  R(C; theta_c) = C_R alpha/(beta pi) ln^2(2C/theta_c^beta)"""
    import gpu_helpers as H
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200.numerics.radiators.generation import gen_crit_sub_num_rad_fused
    grid = O.lin_log_mixed_list(1e-6, 1., 16)
    igs = [Integrand.radiator('log', 1e-10, beta=2., jet_type='gluon', fixedcoupling=False, obs_acc='MLL',
                              split_acc='MLL', theta_scale=float(t)) for t in grid]
    n, seed = 2_000_000, 12
    edges = np.array([np.logspace(-14, -1, 60) * t ** 2 for t in grid])
    res = fused.integrate_batch(igs, n, edges=edges, seed=seed, precision='f32', last_bc=[0., 'plus'])
    for i in (0, 5, len(grid) - 1):
        r = H.run(igs[i], n, seed, edges[i], 'f32')
        assert_array_equal(res['count'][i], r['count'].cpu().numpy())
        one = fused.integrate(igs[i], n, edges=edges[i], seed=seed, precision='f32', last_bc=[0., 'plus'])
        assert_allclose(res['density'][i], one.density, rtol=1e-6, atol=1e-9 * np.abs(one.density).max())
        assert_allclose(res['integral'][i], np.asarray(one.integral), rtol=1e-6, atol=1e-9)
    # data-dependent bins (two passes) against the single-job path in FP64 binning mode
    res2 = fused.integrate_batch(igs, n, numbins=40, binspacing='log', min_log_bins=grid ** 2. * 1e-15, seed=seed,
                                 precision='f32', last_bc=[0., 'plus'])
    assert res2['bins'].shape == (len(grid), 40) and np.all(np.diff(res2['bins'], axis=1) > 0)
    for i in (2, 9):
        one = fused.integrate(igs[i], n, numbins=40, binspacing='log', min_log_bin=grid[i] ** 2. * 1e-15, seed=seed,
                              precision='f32', last_bc=[0., 'plus'])
        assert_allclose(res2['bins'][i], one.bins, rtol=1e-12)
        assert abs(int(res2['count'][i].sum()) - int(one._count.sum().item())) <= 2e-5 * n
        assert_allclose(res2['integral'][i], np.asarray(one.integral), rtol=1e-3, atol=1e-5)
    # physics: fixed coupling LL radiator on the grid
    fn, d = gen_crit_sub_num_rad_fused(4_000_000, jet_type='quark', beta=2., epsilon=1e-8, num_bins=24, seed=3)
    xs, ys, integral = (d[k].reshape(len(set(d['ys'])), -1) for k in ('xs', 'ys', 'integral'))
    exact = O.rad_sub_fc(xs, 2., 'quark', max_radius=ys)
    sel = (xs > 1e-6 * ys ** 2) & (xs < .4 * ys ** 2)
    err = np.abs(integral - exact)[sel] / (exact[sel] + 1e-3)
    assert err.max() < .03, err.max()


@pytest.mark.parametrize("kind,jet,fc,acc,zc,beta,z_pre,f,radius", [
    ('crit', 'quark', False, 'MLL', .05, 3., .02, .3, 1.),
    ('crit', 'gluon', True, 'LL', .2, 1.5, 0., 1., .7),
    ('crit_sub', 'gluon', False, 'MLL', .05, 3., None, 0., 1.),
    ('crit_sub', 'quark', False, 'LL', .2, 1.5, .05, .5, 1.),
    ('crit_sub', 'quark', True, 'MLL', .2, 4., None, 0., .8),
    ('radiator', 'gluon', False, 'MLL', .1, 3., 0., 1., .6),
    ('radiator', 'quark', False, 'LL_nonsing', .1, 1.5, 0., 1., 1.),
])
def test_fast_per_sample_other_parameters(cuda, kind, jet, fc, acc, zc, beta, z_pre, f, radius):
    """the FP32 path away from the headline parameters: z_cut, beta (incl. non-integer), partial grooming
    (z_pre, f), jet radius, LL_nonsing splitting -- per sample against the FP64 reference formulas"""
    import gpu_helpers as H
    from jetmontecarlo_b200.fused import Integrand
    eps = 1e-8
    obs_acc = 'LL' if acc == 'LL_nonsing' else acc
    if kind == 'radiator':
        ig = Integrand.radiator('log', eps, radius=radius, beta=beta, jet_type=jet, fixedcoupling=fc, obs_acc=obs_acc,
                                split_acc=acc, nan_to_num=True)
    elif kind == 'crit':
        ig = Integrand.critical('log', zc, eps, radius=radius, beta=beta, jet_type=jet, fixedcoupling=fc,
                                obs_acc=obs_acc, z_pre=z_pre, f=f, nan_to_num=True)
    else:
        ig = Integrand.crit_sub('log', zc, eps, radius=radius, beta=beta, jet_type=jet, fixedcoupling=fc,
                                obs_acc=obs_acc, z_pre=z_pre, f=f)
    zp = zc if z_pre is None else z_pre
    n = 300000
    s, obs, w = H.dump(ig, n, 99, 'f32')
    z, thc, zs, ths, delta = s[:, 0], s[:, 1], s[:, 2], s[:, 3], s[:, 6]
    with np.errstate(all='ignore'):
        if kind == 'radiator':
            obs_ref = O.ecf_ungroomed(zs, ths, beta, obs_acc)
            w_ref = O.weight_radiator(zs, ths, jet, fc, acc) * zs * ths
            assert ths.max() <= radius * (1 + 1e-6)
        else:
            zcl = zc - zp
            a = delta + (zc - f * zcl)                                  # zmin - f (z_cut - z_pre), from the sampled delta
            b = (1. - z - (1. - f) * zcl) if obs_acc == 'MLL' else (1. - (1. - f) * zcl)
            ccrit = a * b * thc ** beta / (1. - zc) ** 2.
            jc = delta * thc
            assert thc.max() <= radius * (1 + 1e-6)
            if kind == 'crit':
                obs_ref, w_ref = ccrit, O.weight_critical(z, thc, zc, jet, fc) * jc
            else:
                csub = zs * ths ** beta * ths ** beta
                obs_ref = np.maximum(ccrit, csub)
                w_ref = O.weight_two_emission(np.stack([z, thc], 1), np.stack([zs, ths], 1), zc, beta, jet, fc) \
                    * jc * zs * ths
    assert_allclose(obs, obs_ref, rtol=4e-5)
    nan_ref = np.isnan(w_ref)
    tiny = np.abs(np.nan_to_num(w_ref)) < 1e-30
    mism = (nan_ref != (w == 0)) & ~((w_ref == 0) & (w == 0)) & ~tiny
    assert mism.sum() <= max(3, 3e-5 * n), mism.sum()
    good = ~nan_ref & (w != 0) & ~tiny
    assert good.sum() > 100
    rel = np.abs(w[good] / w_ref[good] - 1)
    # bulk agreement; the rare outliers sit next to zeros of the weight (see test_fast_per_sample_vs_oracle)
    assert np.quantile(rel, .999) < 3e-4 and np.median(rel) < 2e-5, (np.quantile(rel, .999), np.median(rel))
    assert abs(w[good].sum() / w_ref[good].sum() - 1) < 1e-4


@pytest.mark.parametrize("jet,zc,beta", [('quark', .1, 2.), ('gluon', .05, 3.), ('quark', .2, 4.)])
def test_critical_sudakov_against_closed_form(cuda, jet, zc, beta):
    """The reference validates its critical-emission integration by plotting it over critSudakov_fc_LL
    (examples/tests/oneEmission_tests/test_fixedcoupcritsudakov.py:124-220).  Here as a number: the closed form is
    the exact cdf of (z - z_c) theta^beta, and C_groomed = that / (1 - z_c)^2, so the Monte Carlo cdf at edge x
    must equal the closed form at x (1 - z_c)^2 -- within 1e-3 absolute at 1e7 samples (north star) and within
    errors at 1e9."""
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    ig = Integrand.critical('log', zc, 1e-10, beta=beta, jet_type=jet, fixedcoupling=True)
    edges = np.logspace(-8, np.log10(.5), 60)
    exact = O.sudakov_crit_fc(edges * (1. - zc) ** 2., zc, beta, jet)
    for n, prec, seed in ((10_000_000, 'f32', 1), (10_000_000, 'f64', 2), (1_000_000_000, 'f32', 3)):
        it = fused.integrate(ig, n, edges=edges, seed=seed, precision=prec, last_bc=[1., 'minus'])
        integ = np.asarray(it.integral)
        # 1e-3 absolute on the cdf, plus the run's own statistical error where that is not yet below 1e-3
        # (log sampling at 1e7 samples: sigma ~ 1e-3 on the cdf of this integrand)
        assert np.abs(integ - exact).max() < 1e-3 + (3. * it.integralErr.max() if n < 10 ** 8 else 0.), \
            (n, prec, np.abs(integ - exact).max())
        pull = (integ[:-1] - exact[:-1]) / (it.integralErr + 1e-7)
        assert np.abs(pull).max() < 4.5, (n, prec, np.abs(pull).max())
