"""Per-event inverse-transform sampling (numerics/sudakov_sampling.py; tensor code that runs on the GPU in the
product) against the per-event loop of the reference's workflow, restated with the oracle's samples_from_cdf
(pinned to the reference's tabulated branches): examples/sudakov_comparisons/sudakov_utils.py:134-238."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose, assert_array_equal
from scipy import interpolate

from oracle import jmc_oracle as O


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


def _tensors(*arrays):
    return [torch.from_numpy(np.asarray(a, dtype=np.float64)) for a in arrays]


def test_probe_grids_match_the_reference_grid():
    from jetmontecarlo_b200.numerics.sudakov_sampling import probe_points
    for lo, hi in ((0., .1), (0., 3.7e-5), (1e-8, 1.), (2.5e-3, .1), (-4., 4.)):
        got = probe_points(*_tensors([lo, lo], [hi, hi])).numpy()
        ref = O.cdf_probe_points([lo, hi])
        assert got.shape == (2, len(ref))
        assert_allclose(got[0], ref, rtol=4e-16, atol=0)        # 10**x of torch and NumPy may differ in the last bit
        assert_allclose(got[0], got[1], rtol=4e-16, atol=0)     # (vectorised pow is not position-independent)


def test_table_inversion_is_interp1d():
    from jetmontecarlo_b200.numerics.sudakov_sampling import invert_tables
    rng = np.random.RandomState(3)
    n, m = 50, 40
    x = np.sort(rng.rand(n, m), axis=1)
    c = np.sort(rng.rand(n, m), axis=1)
    c[:, 5] = c[:, 6]                                             # a plateau
    c[::2] = c[::2, rng.permutation(m)]                           # unsorted rows: interp1d sorts by the abscissa
    keep = rng.rand(n, m) > .2
    keep[:, :2] = True
    r = rng.rand(n) * 1.4 - .2                                    # some beyond the table: extrapolation
    got = invert_tables(*_tensors(c, x), torch.from_numpy(keep), _tensors(r)[0]).numpy()
    ref = np.array([interpolate.interp1d(c[i][keep[i]], x[i][keep[i]], fill_value='extrapolate')(r[i]) for i in range(n)])
    assert_array_equal(got, ref)


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


@pytest.mark.parametrize("force_monotone", [True])
def test_precritical_samples_equal_the_per_event_loop(force_monotone):
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    zs, ths, R = _pre_radiator_table()
    rgi = interpolate.RegularGridInterpolator((zs, ths), R, method='linear', bounds_error=False, fill_value=None)
    rng = np.random.RandomState(5)
    n, z_cut = 400, .1
    theta = np.concatenate([[0., 1e-7, 1.], np.exp(rng.uniform(np.log(1e-6), 0., n - 3))])
    r = rng.rand(n)

    def cdf_np(z, th):
        z = np.atleast_1d(z)
        return np.exp(-1. * rgi(np.stack([z, np.full_like(z, th)], axis=-1)))

    ref = _loop(cdf_np, theta, np.zeros(n), np.full(n, z_cut), r, force_monotone=force_monotone)
    rad = S.grid_interpolant(zs, ths, R)
    got, w = S.generate_precritical_samples(rad, _tensors(theta)[0], z_cut, uniforms=_tensors(r)[0])
    got, w = got.numpy(), w.numpy()
    assert np.all(w == 1.)
    # theta = 0: the radiator vanishes, the cdf table is constant 1 and "monotone", and inverting a constant table
    # divides by zero -- NaN in the reference too (its "cdf always 1" zero-bin test sits behind the monotone test)
    assert np.isnan(ref[0]) and np.isnan(got[0]) and (ref[1:] > 0).all()
    close = np.isclose(got, ref, rtol=1e-9, atol=1e-300, equal_nan=True)
    assert close.mean() > .99, close.mean()
    # (the handful that differ sit on a monotonicity test decided by the last bit of the interpolated radiator:
    #  1e-16 differences between two bilinear formulas flip "cdf[k] <= cdf[k+1]" on a plateau)
    assert_allclose(got[close], ref[close], rtol=1e-9, atol=1e-300)
    assert np.all((got[1:] >= 0) & (got[1:] <= z_cut * 1.001))
    print("pre-critical: %d of %d events identical to the loop, %d within 1e-9" % ((got == ref).sum(), n, close.sum()))


def test_subsequent_samples_and_every_branch():
    """synthetic conditional cdfs that hit each shape of the list, with per-event domains [0, theta^beta / 2]"""
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    beta = 2.
    rng = np.random.RandomState(8)
    n = 300
    theta = np.concatenate([[1e-6, 1e-5], np.exp(rng.uniform(np.log(2e-5), 0., n - 2))])
    r = rng.rand(n)

    def rad_np(c, th):
        cmax = th ** beta / 2.
        with np.errstate(all='ignore'):
            base = 1.3 * np.log(cmax / np.maximum(c, 1e-300)) ** 2 * (c < cmax) * (c > 0) + 5000. * (c <= 0)
        wiggle = .05 * np.sin(40. * np.log(np.maximum(c, 1e-300))) * (th > .1) * (c > 0)      # breaks monotonicity
        head = 400. * (c < 1e-4 * cmax) * (th > .5)                                          # negligible head
        return base * (1. + wiggle) + head

    def rad_t(c, th):
        cmax = th ** beta / 2.
        tiny = torch.full_like(c, 1e-300)
        base = 1.3 * torch.log(cmax / torch.maximum(c, tiny)) ** 2 * (c < cmax) * (c > 0) + 5000. * (c <= 0)
        wiggle = .05 * torch.sin(40. * torch.log(torch.maximum(c, tiny))) * (th > .1) * (c > 0)
        head = 400. * (c < 1e-4 * cmax) * (th > .5)
        return base * (1. + wiggle) + head

    upper = theta ** beta / 2.
    live = ~(upper < 1e-10)
    ref = np.full(n, 1e-100)
    ref[live] = _loop(lambda c, th: np.exp(-1. * rad_np(np.atleast_1d(c), th)), theta[live], np.zeros(live.sum()),
                      upper[live], r[live], force_monotone=True)
    got, w = S.generate_subsequent_samples(rad_t, _tensors(theta)[0], beta, uniforms=_tensors(r)[0])
    got = got.numpy()
    assert (~live).sum() == 2 and np.all(got[~live] == 1e-100)
    close = np.isclose(got, ref, rtol=1e-8, atol=1e-300)
    assert close.mean() > .98, close.mean()
    assert np.all(got[live] <= upper[live] * 1.01) and np.all(w.numpy() == 1.)

    # without force_monotone a wiggling cdf is refused, naming the event
    with pytest.raises(ValueError, match="is not monotone"):
        S.conditional_samples_from_cdf(lambda c, th: torch.exp(-rad_t(c, th)), *_tensors(theta[live], np.zeros(live.sum()),
                                       upper[live], r[live]))
    # zero bin: a cdf that is 1 at the lower end of the domain and drops right after it
    conds = np.array([.5, 2.])
    got, _ = S.conditional_samples_from_cdf(lambda x, c: torch.where(x == 0, torch.ones_like(x), .5 + .5 * x / c),
                                            *_tensors(conds, np.zeros(2), conds, np.array([.3, .6])))
    ref = _loop(lambda x, c: np.where(x == 0, 1., .5 + .5 * x / c), conds, np.zeros(2), conds, np.array([.3, .6]))
    assert_array_equal(got.numpy(), ref) and np.all(ref == 0.)
    # overflow bin: negligible everywhere (with a dip, so that the table is not trivially monotone)
    got, _ = S.conditional_samples_from_cdf(lambda x, c: 1e-30 * (2. + torch.sin(50 * x / c)),
                                            *_tensors(conds, np.zeros(2), conds, np.array([.3, .6])))
    assert_array_equal(got.numpy(), conds)
    # catch_turnpoint: a cdf that reaches 1 at 60 % of its domain and turns back
    def turning(x, cond):
        frac = x / cond
        return torch.where(frac < .6, torch.clamp(frac / .4, max=1.), 1. - (frac - .6))

    conds = np.array([.5, 1., 2.])
    rr = np.array([.25, .5, .75])
    got, _ = S.conditional_samples_from_cdf(turning, *_tensors(conds, np.full(3, 1e-6), conds, rr), catch_turnpoint=True)
    ref = _loop(lambda x, c: np.where(x / c < .6, np.minimum(x / c / .4, 1.), 1. - (x / c - .6)), conds, np.full(3, 1e-6),
                conds, rr, catch_turnpoint=True)
    assert_allclose(got.numpy(), ref, rtol=1e-10)


def test_grid_interpolant_is_regular_grid_interpolator():
    from jetmontecarlo_b200.numerics.sudakov_sampling import grid_interpolant
    zs, ths, R = _pre_radiator_table()
    rgi = interpolate.RegularGridInterpolator((zs, ths), R, method='linear', bounds_error=False, fill_value=None)
    rng = np.random.RandomState(2)
    k = min(len(zs[::7]), len(ths[::7]))
    x = np.concatenate([zs[::7][:k], rng.uniform(-.01, .15, 500)])
    y = np.concatenate([ths[::7][:k], rng.uniform(-.1, 1.2, 500)])
    got = grid_interpolant(zs, ths, R)(*_tensors(x, y)).numpy()
    assert_allclose(got, rgi(np.stack([x, y], axis=-1)), rtol=1e-12, atol=1e-12)
