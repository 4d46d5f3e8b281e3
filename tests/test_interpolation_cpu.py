"""Host-side interpolation helpers against vectors produced by the reference's own functions
(oracle/make_golden.py::interpolation_helpers; ref: utils/interpolation.py:43-273)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal


def test_1d_interpolants_are_the_reference(golden):
    from jetmontecarlo_b200.utils import interpolation as I
    g = golden("interpolation")
    for name in ('rad_log', 'lin_inc', 'wiggle'):
        xs, fs, probe = g[f'{name}:xs'], g[f'{name}:fs'], g[f'{name}:probe']
        if name != 'wiggle':
            assert_array_equal(I.get_1d_interpolation(xs, fs, monotonic=True)(probe), g[f'{name}:mono_default'])
            assert_array_equal(I.get_1d_interpolation(xs, fs, monotonic=True, bounds=(xs[2], xs[-3]),
                                                      bound_values=[None, None])(probe), g[f'{name}:mono_bounds'])
        assert_array_equal(I.get_1d_interpolation(xs, fs, monotonic=False, bounds=(xs[0], xs[-1]),
                                                  bound_values=[0., 0.])(probe), g[f'{name}:free_bounds'])
        assert_array_equal(I.get_1d_interpolation(xs, fs, monotonic=False, bounds=(xs[1], xs[-2]),
                                                  bound_values=None)(probe), g[f'{name}:free_cont'])
        assert_array_equal([I.is_monotonic_arr(fs), I.is_monotonic_arr(fs, 'increasing'),
                            I.is_monotonic_arr(fs, 'decreasing')], g[f'{name}:is_mono'])
        assert_array_equal(I.where_monotonic_arr(fs, 'increasing'), g[f'{name}:where_inc'])
        assert_array_equal(I.where_monotonic_arr(fs, 'decreasing'), g[f'{name}:where_dec'])
    xs, fs = g['wiggle:xs'], g['wiggle:fs']
    with pytest.raises(AssertionError, match="not monotone"):
        I.get_1d_interpolation(xs, np.sin(7. * xs), monotonic=True)
    with pytest.raises(AssertionError, match="Cannot assign boundary values"):
        I.get_1d_interpolation(xs, fs, monotonic=False)
    with pytest.raises(AssertionError, match="length 2"):
        I.get_1d_interpolation(xs, fs, monotonic=False, bounds=(0., 1.), bound_values=(0.,))
    # the documented first-bin behaviour of the default monotone interpolant, and the way around it
    xs, fs = g['lin_inc:xs'], g['lin_inc:fs']
    default = I.get_1d_interpolation(xs, fs, monotonic=True)
    assert default(xs[5]) == fs[-1] and default(.5 * (xs[0] + xs[1])) == .5 * (fs[0] + fs[1])
    whole = I.get_1d_interpolation(xs, fs, monotonic=True, bounds=(xs[0], xs[-1]))
    assert_array_equal(whole(xs[1:-1]), fs[1:-1])
    # no state leaks from one call to the next (the reference's default list does)
    first = I.get_1d_interpolation(g['rad_log:xs'], g['rad_log:fs'], monotonic=True)(g['rad_log:probe'])
    assert_array_equal(I.get_1d_interpolation(xs, fs, monotonic=True)(g['lin_inc:probe']), g['lin_inc:mono_default'])
    assert_array_equal(first, g['rad_log:mono_default'])


def test_monotonic_domain_and_func(golden):
    from jetmontecarlo_b200.utils import interpolation as I
    g = golden("interpolation")
    funcs = {'square': lambda x: x ** 2, 'sin': lambda x: np.sin(3. * x), 'cubic': lambda x: x ** 3 - x}
    cases = [('square', [-10, 10], 1, 'increasing'), ('square', [0, 10], 1, 'increasing'),
             ('square', [-10, 10], 1, 'decreasing'), ('square', [-10, 10], -1, 'decreasing'),
             ('square', [-10, 10], -1, 'increasing'), ('sin', [0.1, 3.], 1.5, None),
             ('sin', [-2., 2.], 0.1, 'increasing'), ('cubic', [-2., 2.], 0., None),
             ('cubic', [.1, 2.], None, None), ('square', [1e-3, 4.], None, 'increasing')]
    got = []
    for fname, domain, point, check in cases:
        res = I.monotonic_domain(funcs[fname], domain, point, check)
        got.append([np.nan, np.nan] if res is None else [float(res[0]), float(res[1])])
    assert_array_equal(np.array(got), g['monotonic_domain'])
    assert_array_equal([I.is_monotonic_func(funcs['square'], [.1, 3.]), I.is_monotonic_func(funcs['sin'], [.1, 3.]),
                        I.is_monotonic_func(funcs['sin'], [.1, .4], 'increasing'),
                        I.is_monotonic_func(funcs['square'], [.1, 3.], 'decreasing')], g['is_monotonic_func'])
    with pytest.raises(AssertionError):
        I.monotonic_domain(funcs['square'], [0, 1], 2.)


def test_2d_interpolants(golden):
    from jetmontecarlo_b200.utils import interpolation as I
    g = golden("interpolation")
    x, y, z, pts = g['grid:x'], g['grid:y'], g['grid:z'], g['grid:pts']
    assert_array_equal(I.get_2d_interpolation(x, y, z)(pts), g['grid:linear'])
    assert_array_equal(I.get_2d_interpolation(x, y, z, method='nearest')(pts), g['grid:nearest_method'])
    xx, yy = np.meshgrid(x, y, indexing='ij')
    near = I.get_2d_interpolation(xx.flatten(), yy.flatten(), z.flatten(), interpolation_method='Nearest')
    assert_array_equal(near(pts[:, 0], pts[:, 1]), g['scatter:nearest'])
    with pytest.raises(ValueError, match="Unknown interpolation method"):
        I.get_2d_interpolation(x, y, z, interpolation_method='spline')
    with pytest.raises(NotImplementedError, match="interp2d"):
        I.get_2d_interpolation(x, y, z, interpolation_method='Cubic')


def test_hist_postprocessing(golden, capsys):
    """histDerivative / smooth_data on a golden cumulative integral, bit for bit (ref: utils/hist_utils.py:7-91, 158-173)"""
    from jetmontecarlo_b200.utils.hist_utils import histDerivative, nan_info, notfinite_info, smooth_data
    g = golden("hist_postprocessing")
    for name, spacing in (('c1_log0', 'log'), ('c1_lin1', 'lin')):
        c1 = golden(name)
        edges, hist = c1['quark_fc_LL:edges'], c1['quark_fc_LL:integral'][:-1]
        interp, deriv = histDerivative(hist, edges, giveHist=True, binInput=spacing)
        assert_array_equal(deriv, g[f'{name}:deriv'])
        assert_array_equal(interp(g[f'{name}:probe']), g[f'{name}:interp'])
        assert_array_equal(histDerivative(hist, edges, binInput=spacing)(g[f'{name}:probe']), g[f'{name}:interp'])
        mids = np.exp((np.log(edges[1:]) + np.log(edges[:-1])) / 2.) if spacing == 'log' else (edges[1:] + edges[:-1]) / 2.
        assert_array_equal(smooth_data(mids, hist), g[f'{name}:smooth'])
    with pytest.raises(AssertionError, match="either 'lin' or 'log'"):
        histDerivative(hist, edges, binInput='sqrt')
    nan_info([1., np.nan, 2.], "w")
    notfinite_info([1., np.nan, np.inf], "w")
    assert capsys.readouterr().out.split() == ['nan', 'w', '1', 'notfinite', 'w', '2']
