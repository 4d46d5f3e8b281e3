"""Per-event inverse-transform sampling on the device (jmc_conditional_inverse_transform: one event per thread block)
against the per-event loop of the reference's workflow, restated with the oracle's samples_from_cdf (pinned to the
reference's tabulated branches) over scipy's RegularGridInterpolator of the same radiator table.
ref: examples/sudakov_comparisons/sudakov_utils.py:134-238, utils/montecarlo_utils.py:142-318, 377-399"""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose, assert_array_equal
from scipy import interpolate

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


def _loop(cdf_np, conds, lows, highs, r, **kw):
    """what the reference does: one samples_from_cdf call per event"""
    out = []
    for i in range(len(conds)):
        with np.errstate(all='ignore'):
            s, w = O.samples_from_cdf_tabulated(lambda x, c=conds[i]: cdf_np(x, c), 1, [lows[i], highs[i]],
                                                lambda n, i=i: r[i:i + 1], **kw)
        assert w[0] == 1.
        out.append(s[0])
    return np.array(out)


def _table_cdf(xs, ys, R):
    rgi = interpolate.RegularGridInterpolator((xs, ys), R, method='linear', bounds_error=False, fill_value=None)

    def cdf(x, cond):
        x = np.atleast_1d(x)
        return np.exp(-1. * rgi(np.stack([x, np.full_like(x, cond)], axis=-1)))
    return cdf


def _pre_radiator_table():
    """a pre-critical-radiator-like table R(z_pre, theta) on a rectangular grid with Monte Carlo noise, so that some
    conditional cdfs are not monotone, some start at exactly 1 and some are negligible at small z_pre"""
    rng = np.random.RandomState(11)
    zs = np.concatenate([[0.], np.logspace(-8, -1, 60)])
    ths = np.concatenate([[0.], np.logspace(-6, 0, 45)])
    Z, T = np.meshgrid(zs, ths, indexing='ij')
    with np.errstate(all='ignore'):
        R = .8 * np.log(.1 / np.maximum(Z, 1e-300)) * np.log(1. / np.maximum(T, 1e-300)) * (T > 0) * (Z > 0)
    R[0, :] = 2000.                                               # cdf underflows to 0 at z_pre = 0 ...
    R[:, 0] = 0.                                                  # ... and is 1 for theta = 0
    R *= 1. + .03 * rng.randn(*R.shape) * (T > 1e-3)
    R[zs < 1e-6] = np.where(ths[None, :] > .2, 900., R[zs < 1e-6])    # a negligible head for large angles
    return zs, ths, R


def test_precritical_samples_equal_the_per_event_loop(cuda):
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    zs, ths, R = _pre_radiator_table()
    cdf_np = _table_cdf(zs, ths, R)
    rng = np.random.RandomState(5)
    n, z_cut = 400, .1
    theta = np.concatenate([[0., 1e-7, 1.], np.exp(rng.uniform(np.log(1e-6), 0., n - 3))])
    r = rng.rand(n)
    ref = _loop(cdf_np, theta, np.zeros(n), np.full(n, z_cut), r, force_monotone=True)
    rad = S.grid_interpolant(zs, ths, R)
    got, w = S.generate_precritical_samples(rad, theta, z_cut, uniforms=torch.from_numpy(r).cuda())
    assert got.is_cuda
    got, w = got.cpu().numpy(), w.cpu().numpy()
    assert np.all(w == 1.)
    # theta = 0: the radiator vanishes, the cdf table is constant 1 and "monotone", and inverting a constant table
    # divides by zero -- NaN in the reference too (its "cdf always 1" zero-bin test sits behind the monotone test)
    assert np.isnan(ref[0]) and np.isnan(got[0]) and (ref[1:] > 0).all()
    close = np.isclose(got, ref, rtol=1e-9, atol=1e-300, equal_nan=True)
    # (the handful that differ sit on a monotonicity test decided by the last bit of the interpolated radiator:
    #  1e-16 differences between two bilinear formulas, or between the device's and libm's exp / pow, flip
    #  "cdf[k] <= cdf[k+1]" on a plateau and move that event's sample)
    assert close.mean() > .97, close.mean()
    assert_allclose(got[close], ref[close], rtol=1e-9, atol=1e-300)
    assert np.all((got[1:] >= 0) & (got[1:] <= z_cut * 1.001))
    # the radiator object evaluates like scipy's interpolant (host convenience)
    pts_x, pts_y = rng.uniform(-.01, .15, 200), rng.uniform(-.1, 1.2, 200)
    rgi = interpolate.RegularGridInterpolator((zs, ths), R, method='linear', bounds_error=False, fill_value=None)
    assert_allclose(rad(pts_x, pts_y), rgi(np.stack([pts_x, pts_y], axis=-1)), rtol=1e-12, atol=1e-12)


def _sub_radiator_table(beta=2.):
    """R(C, theta) on a grid: quadratic-log radiator inside C < theta^beta / 2, with a wiggle that breaks
    monotonicity at large angles and a plateau that makes the head of the cdf negligible there"""
    cs = np.concatenate([[0.], np.logspace(-14, np.log10(.5), 140)])
    ths = np.logspace(-5, 0, 70)
    Cg, Tg = np.meshgrid(cs, ths, indexing='ij')
    cmax = Tg ** beta / 2.
    with np.errstate(all='ignore'):
        base = 1.3 * np.log(cmax / np.maximum(Cg, 1e-300)) ** 2 * (Cg < cmax) * (Cg > 0) + 5000. * (Cg <= 0)
        wiggle = .05 * np.sin(40. * np.log(np.maximum(Cg, 1e-300))) * (Tg > .1) * (Cg > 0)
    head = 400. * (Cg < 1e-4 * cmax) * (Tg > .5)
    return cs, ths, base * (1. + wiggle) + head


def test_subsequent_samples_and_domains_per_event(cuda):
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    beta = 2.
    cs, ths, R = _sub_radiator_table(beta)
    cdf_np = _table_cdf(cs, ths, R)
    rng = np.random.RandomState(8)
    n = 300
    theta = np.concatenate([[1e-6, 1e-5], np.exp(rng.uniform(np.log(2e-5), 0., n - 2))])
    r = rng.rand(n)
    upper = theta ** beta / 2.
    live = ~(upper < 1e-10)
    ref = np.full(n, 1e-100)
    ref[live] = _loop(cdf_np, theta[live], np.zeros(live.sum()), upper[live], r[live], force_monotone=True)
    rad = S.grid_interpolant(cs, ths, R)
    got, w = S.generate_subsequent_samples(rad, theta, beta, uniforms=r)
    got = got.cpu().numpy()
    assert (~live).sum() == 2 and np.all(got[~live] == 1e-100)
    close = np.isclose(got, ref, rtol=1e-8, atol=1e-300)
    assert close.mean() > .97, close.mean()
    # (a forced-monotone table drops entries: a uniform above the remaining cdf values extrapolates past the domain,
    #  in the reference as here)
    assert np.all((got[live] <= upper[live] * 1.01) | (ref[live] > upper[live] * 1.01) | ~close[live])
    assert np.all(w.cpu().numpy() == 1.)
    # without force_monotone a wiggling cdf is refused
    th = torch.from_numpy(theta[live]).cuda()
    with pytest.raises(ValueError, match="is not monotone"):
        S.conditional_samples_from_table(rad, th, torch.zeros_like(th), th ** beta / 2., torch.from_numpy(r[live]).cuda())


def _one_d_table(fn, lo, hi, conds, npts=4001):
    """a table whose rows are an arbitrary cdf shape: R(x, cond) = -log(fn(x / cond)) sampled densely, so that linear
    interpolation in between reproduces the shape closely enough to land in the same branch"""
    xs = np.linspace(lo, hi, npts)
    ys = np.asarray(conds, dtype=float)
    with np.errstate(all='ignore'):
        R = -np.log(np.stack([fn(xs, c) for c in ys], axis=1))
    return xs, ys, R


def test_every_branch_of_the_shape_list(cuda):
    """zero bin, overflow bin, negligible head, turning point: tables built to hit each branch, device vs the oracle's
    walk over scipy's interpolant of the same table"""
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    conds = np.array([.5, 2.])
    dev = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).cuda()

    def run(xs, ys, R, lo, hi, rr, **kw):
        rad = S.TableRadiator(xs, ys, R)
        got, _ = S.conditional_samples_from_table(rad, dev(ys), dev(lo), dev(hi), dev(rr), **kw)
        ref = _loop(_table_cdf(xs, ys, R), ys, lo, hi, rr, **kw)
        return got.cpu().numpy(), ref

    # zero bin: a cdf that is 1 at the lower end of the domain and lower right after it (a radiator table that jumps
    # from 0 at x = 0 to a finite value one tiny grid step later, then falls back to 0 at the end of the domain)
    xs = np.concatenate([[0., 1e-12], np.logspace(-9, np.log10(2.), 50)])
    R = np.stack([np.concatenate([[0., 5.], np.linspace(5., 0., 50)])] * 2, axis=1)
    got, ref = run(xs, conds, R, np.zeros(2), np.full(2, 2.), np.array([.3, .6]))
    assert_array_equal(got, ref)
    assert np.all(ref == 0.)
    # overflow bin: negligible everywhere (with a dip, so that the table is not trivially monotone)
    xs, ys, R = _one_d_table(lambda x, c: 1e-30 * (2. + np.sin(50 * x / c)), 0., 2., conds)
    got, ref = run(xs, ys, R, np.zeros(2), conds, np.array([.3, .6]))
    assert_array_equal(got, conds)
    assert_array_equal(ref, conds)
    # negligible head, monotone afterwards: the domain is cut and the walk starts again
    xs, ys, R = _one_d_table(lambda x, c: np.where(x < .3 * c, 1e-25 * (1.5 + np.cos(90 * x / c)), (x / c - .3) / .7 + 1e-12),
                             0., 2., conds)
    got, ref = run(xs, ys, R, np.zeros(2), conds, np.array([.25, .8]))
    assert_allclose(got, ref, rtol=1e-9)
    assert np.all(got > .3 * conds)
    # catch_turnpoint: a cdf that reaches 1 at 60 % of its domain and turns back
    conds3 = np.array([.5, 1., 2.])
    turning = lambda x, c: np.where(x / c < .6, np.minimum(x / c / .4, 1.), 1. - (x / c - .6))
    xs, ys, R = _one_d_table(turning, 1e-6, 2., conds3)
    got, ref = run(xs, ys, R, np.full(3, 1e-6), conds3, np.array([.25, .5, .75]), catch_turnpoint=True)
    assert_allclose(got, ref, rtol=1e-9)
    # nearest-node interpolation of the table (get_2d_interpolation's 'Nearest')
    zs, ths, Rp = _pre_radiator_table()
    rad = S.TableRadiator(zs, ths, Rp, method="Nearest")
    th = np.exp(np.random.RandomState(2).uniform(np.log(1e-5), 0., 64))
    got, _ = S.conditional_samples_from_table(rad, dev(th), dev(np.zeros(64)), dev(np.full(64, .1)),
                                              dev(np.random.RandomState(3).rand(64)), force_monotone=True)
    ref = _loop(lambda x, c: np.exp(-1. * rad(np.atleast_1d(x), c)), th, np.zeros(64), np.full(64, .1),
                np.random.RandomState(3).rand(64), force_monotone=True)
    close = np.isclose(got.cpu().numpy(), ref, rtol=1e-9, atol=1e-300)
    assert close.mean() > .9, close.mean()


def test_large_event_counts_and_counter_rng(cuda):
    """20000 events with the uniforms of the counter RNG: distribution-level agreement with a per-event loop on a
    subset, bounds, and run-to-run reproducibility"""
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    zs, ths, R = _pre_radiator_table()
    n, seed, z_cut = 20000, 9, .1
    theta = np.exp(np.random.RandomState(1).uniform(np.log(1e-6), 0., n))
    rad = S.grid_interpolant(zs, ths, R)
    got, w = S.generate_precritical_samples(rad, theta, z_cut, seed=seed)
    again, _ = S.generate_precritical_samples(rad, theta, z_cut, seed=seed)
    assert got.is_cuda and bool((w == 1.).all()) and torch.equal(got, again)
    got = got.cpu().numpy()
    assert np.all((got >= 0) & (got <= z_cut * 1.001))
    r = O.philox_uniforms53(seed, 0, n, 1)[:, 0]
    sub = slice(0, 300)
    ref = _loop(_table_cdf(zs, ths, R), theta[sub], np.zeros(300), np.full(300, z_cut), r[sub], force_monotone=True)
    assert np.isclose(got[sub], ref, rtol=1e-9, atol=1e-300).mean() > .97


def test_generators_edge_cases(cuda):
    """empty input, the underflow rule, a zero critical angle, NULL arguments: jmc_sudakov_generate itself"""
    import ctypes as C
    import torch
    from jetmontecarlo_b200 import _device as D, _lib
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    gx, gy = np.logspace(-10, -1, 30), np.logspace(-8, 0, 25)
    table = np.outer(np.linspace(4., 0., 30), np.linspace(2., .5, 25))
    rad = S.TableRadiator(gx, gy, table)
    s, w = S.generate_precritical_samples(rad, np.array([]), .1, seed=1)
    assert s.numel() == 0 and w.numel() == 0
    # subsequent: theta^beta / 2 < 1e-10 -> 1e-100 with weight 1, without sampling
    theta = np.array([1e-6, 1e-5 * .9, .3, 0.])
    s, w = S.generate_subsequent_samples(S.TableRadiator(gx, gy, table, "Nearest"), theta, 2., uniforms=np.full(4, .5))
    s, w = s.cpu().numpy(), w.cpu().numpy()
    assert s[0] == 1e-100 and s[1] == 1e-100 and s[3] == 1e-100 and np.all(w == 1.)
    assert 0. <= s[2] <= .3 ** 2 / 2.
    # the same uniforms give the same samples whether passed in or drawn from the event's Philox stream
    th = np.full(1000, .2)
    u = np.empty((1000, 1))
    ud = D.empty((1000, 1))
    _lib.check(_lib.lib.jmc_uniforms(1000, 77, 0, 1, D.ptr(ud), D.stream()))
    a, _ = S.generate_precritical_samples(rad, th, .1, seed=77)
    b, _ = S.generate_precritical_samples(rad, th, .1, uniforms=ud[:, 0].contiguous())
    np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
    L, null = _lib.lib, C.c_void_p(0)
    dx, dy, dt = rad.device()
    out = D.empty((4,))
    assert L.jmc_sudakov_generate(3, D.ptr(dx), 30, D.ptr(dy), 25, D.ptr(dt), 0, D.ptr(out), .1, null, 1, 4, D.ptr(out),
                                  D.ptr(out), D.stream()) == -1                  # unknown emission
    assert L.jmc_sudakov_generate(1, D.ptr(dx), 30, D.ptr(dy), 25, D.ptr(dt), 0, null, .1, null, 1, 4, D.ptr(out),
                                  D.ptr(out), D.stream()) == -1                  # NULL critical angles
    assert L.jmc_sudakov_generate(1, D.ptr(dx), 1, D.ptr(dy), 25, D.ptr(dt), 0, D.ptr(out), .1, null, 1, 4, D.ptr(out),
                                  D.ptr(out), D.stream()) == -1                  # a 1-point grid axis
    assert L.jmc_sudakov_generate(1, D.ptr(dx), 30, D.ptr(dy), 25, D.ptr(dt), 0, null, .1, null, 1, 0, null, null,
                                  D.stream()) == 0                               # nothing to do
