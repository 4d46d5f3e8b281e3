"""Host-only logic of the integrator mirror (no device call is reached): bin edges from a range, boundary
conditions, state flags and assertion messages, file formats, interpolation.  ref: montecarlo/integrator.py"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal


def _integrator():
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    return integrator()


@pytest.mark.parametrize("name,method", [("c1_log0", "log"), ("c1_log2", "log"), ("c1_lin1", "lin")])
def test_edges_from_range_are_the_reference_bins(golden, name, method):
    """setBins = min/max (device) + this host formula; the formula alone must give the reference's edges and
    midpoints bit for bit from the reference's own min/max.  ref: integrator.py:37-57"""
    g = golden(name)
    for tag in ("quark_fc_LL:", "gluon_rc_MLL:"):
        obs, edges = g[tag + 'obs'], g[tag + 'edges']
        it = _integrator()
        it._edges_from_range(len(edges), obs.min(), obs.max(), method, 1e-15)
        assert_array_equal(it.bins, edges)
        assert_array_equal(it.bin_midpoints, g[tag + 'midpoints'])
        assert it.hasBins and it.binspacing == method


def test_log_bins_floor_and_lin_bins_include_zero():
    it = _integrator()
    it._edges_from_range(5, 0., 10., 'log', 1e-15)          # min obs 0 -> floor at min_log_bin
    assert_allclose(it.bins[0], 1e-15, rtol=1e-12)
    it._edges_from_range(5, 1e-30, 10., 'log', 1e-3)
    assert_allclose(it.bins, np.logspace(-3, 1, 5), rtol=1e-14)
    it._edges_from_range(5, 2., 10., 'lin', 1e-15)          # lin bins always start at min(0, min obs)
    assert_array_equal(it.bins, np.linspace(0., 10., 5))
    it._edges_from_range(5, -2., 10., 'lin', 1e-15)
    assert_array_equal(it.bins, np.linspace(-2., 10., 5))


def test_boundary_condition_arguments():
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.montecarlo.integrator import _bc_args
    assert _bc_args(0.5, None) == (_lib.BC_FIRST, 0.5)
    assert _bc_args(None, [1., 'plus']) == (_lib.BC_LAST_PLUS, 1.)
    assert _bc_args(None, [0., 'minus']) == (_lib.BC_LAST_MINUS, 0.)
    for first, last in ((None, None), (0., [1., 'plus'])):
        with pytest.raises(AssertionError, match="unambiguous boundary conditions"):
            _bc_args(first, last)
    with pytest.raises(AssertionError, match="plus or minus"):
        _bc_args(None, [1., 'sideways'])


def test_state_flags_and_assertion_messages(tmp_path):
    """the reference's flags start False and its assertions fire before any work (integrator.py:367-386, 92-95, 138)"""
    it = _integrator()
    for flag in ('hasBins', 'hasMCDensity', 'hasMCIntegral', 'hasInterpIntegral', 'hasInterpDensity',
                 'useFirstBinBndCond', 'useLastBinBndCond', 'monotone'):
        assert getattr(it, flag) is False
    assert it.firstBinBndCond is None and it.lastBinBndCond is None
    with pytest.raises(AssertionError, match="Need bins to produce density histogram."):
        it.setDensity(np.ones(3), np.ones(3), 1.)
    it.bins, it.hasBins = np.linspace(0, 1, 4), True
    with pytest.raises(AssertionError, match="Need a valid area"):
        it.setDensity(np.ones(3), np.ones(3), None)
    with pytest.raises(AssertionError, match="Need density function to perform integration"):
        it.integrate()
    with pytest.raises(AssertionError, match="Need MC integral to produce interpolation"):
        it.makeInterpolatingFn()
    with pytest.raises(AssertionError, match="No integral histogram to save"):
        it.saveIntegral(*(str(tmp_path / n) for n in "abc"))
    with pytest.raises(AssertionError, match="No density histogram to save"):
        it.saveDensity(*(str(tmp_path / n) for n in "abc"))
    with pytest.raises(AssertionError, match="Need density function and integral to save MC data"):
        it.save_montecarlo_data(str(tmp_path / "x.npz"))
    it.setLastBinBndCondition([1., 'minus'])
    assert it.lastBinBndCond == [1., 'minus'] and it.useLastBinBndCond
    it.setFirstBinBndCondition(0.)
    assert it.firstBinBndCond == 0. and it.useFirstBinBndCond
    it.setMonotone()
    assert it.monotone is True


def test_file_formats_round_trip(tmp_path, golden):
    """csv triplets and the npz keys of the reference (integrator.py:204-224, 320-354), filled from a golden result"""
    g = golden("c1_log0")
    tag = "quark_fc_LL:"
    it = _integrator()
    it.bins, it.bin_midpoints = g[tag + 'edges'], g[tag + 'midpoints']
    it.density, it.densityErr = g[tag + 'density'], g[tag + 'density_err']
    it.integral, it.integralErr = list(g[tag + 'integral']), g[tag + 'integral_err']
    it.hasBins = it.hasMCDensity = it.hasMCIntegral = True
    files = [str(tmp_path / n) for n in ("val.csv", "err.csv", "bins.csv")]
    it.saveIntegral(*files)
    other = _integrator()
    other.loadIntegral(*files)
    assert other.hasMCIntegral
    assert_array_equal(other.integral, g[tag + 'integral'])          # savetxt's %.18e round-trips float64
    assert_array_equal(other.integralErr, g[tag + 'integral_err'])
    assert_array_equal(other.bins, g[tag + 'edges'])
    it.saveDensity(*files)
    other.loadDensity(*files)
    assert other.hasMCDensity
    assert_array_equal(other.density, g[tag + 'density'])
    it.save_montecarlo_data(str(tmp_path / "mc.npz"), info="C1")
    z = np.load(str(tmp_path / "mc.npz"))
    assert set(z.files) == {'bins', 'bin midpoints', 'density', 'density error', 'integral', 'integral error', 'info'}
    assert_array_equal(z['integral'], g[tag + 'integral'])
    assert str(z['info']) == "C1"


def test_interpolating_function_of_a_golden_integral(golden):
    """makeInterpolatingFn on a reference result, as the radiator generators use it (setMonotone first,
    generation.py:84,128); semantics of utils/interpolation.py, see tests/test_interpolation_cpu.py"""
    from jetmontecarlo_b200.utils.interpolation import get_1d_interpolation
    g = golden("c1_log0")
    tag = "quark_fc_LL:"
    it = _integrator()
    it.bins, it.integral, it.hasMCIntegral = g[tag + 'edges'], list(g[tag + 'integral']), True
    # without setMonotone and without bounds the reference refuses (interpolation.py:243-244)
    with pytest.raises(AssertionError, match="Cannot assign boundary values"):
        it.makeInterpolatingFn()
    it.setMonotone()
    it.makeInterpolatingFn()
    assert it.hasInterpIntegral
    probe = np.concatenate([it.bins, np.sqrt(it.bins[1:] * it.bins[:-1])])
    assert_array_equal(it.interpFn(probe), get_1d_interpolation(it.bins, it.integral, monotonic=True)(probe))
    # over the whole table: exact at the interior edges, between the neighbouring values inside a bin
    it.makeInterpolatingFn(bounds=(it.bins[0], it.bins[-1]))
    assert_allclose(it.interpFn(it.bins[1:-1]), g[tag + 'integral'][1:-1], rtol=1e-15)
    mid = np.sqrt(it.bins[1:] * it.bins[:-1])
    v = it.interpFn(mid)
    lo = np.minimum(g[tag + 'integral'][1:], g[tag + 'integral'][:-1])
    hi = np.maximum(g[tag + 'integral'][1:], g[tag + 'integral'][:-1])
    assert np.all(v >= lo - 1e-12) and np.all(v <= hi + 1e-12)


def test_sampler_file_format_and_descriptor(tmp_path, golden):
    """npz layout of the reference (keys 'samples', 'area'; sampler.py:47-60) and the C descriptor of each sampler"""
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.montecarlo.sampler import simpleSampler
    from jetmontecarlo_b200.numerics.radiators.samplers import (criticalSampler, precriticalSampler,
                                                                ungroomedSampler)
    g = golden("crit_log_0.1")
    s = criticalSampler('log', zc=.1, epsilon=float(g['epsilon']))
    s.samples = g['samples']
    f = str(tmp_path / "crit.npz")
    s.saveSamples(f)
    z = np.load(f)
    assert set(z.files) == {'samples', 'area'} and float(z['area']) == float(g['area'])
    t = criticalSampler('log', zc=.1, epsilon=float(g['epsilon']))
    t.loadSamples(f)
    assert_array_equal(t.getSamples(), g['samples'])
    assert float(t.area) == s.area
    t.clearSamples()
    assert t.samples == []
    d = s._descriptor()
    assert (d.kind, d.method, d.zc, d.radius, d.epsilon) == (_lib.SAMPLER_CRITICAL, _lib.LOG, .1, 1., float(g['epsilon']))
    d = precriticalSampler('lin', zc=.2)._descriptor()
    assert (d.kind, d.method, d.zc, d.epsilon) == (_lib.SAMPLER_PRECRITICAL, _lib.LIN, .2, 0.)
    d = ungroomedSampler('log', epsilon=1e-5, radius=.4)._descriptor()
    assert (d.kind, d.method, d.radius, d.epsilon) == (_lib.SAMPLER_UNGROOMED, _lib.LOG, .4, 1e-5)
    d = simpleSampler('lin', bounds=[.25, .75])._descriptor()
    assert (d.kind, d.lo, d.hi) == (_lib.SAMPLER_SIMPLE, .25, .75)


def test_pickled_interpolants_and_analytic_hooks(golden, tmp_path):
    """dill files of the interpolating functions (integrator.py:241-285) and the analytic-integral hooks (:290-315)"""
    from jetmontecarlo_b200.montecarlo.integrator import integrator, multidim_cumsum
    g = golden("c1_lin1")
    tag = "quark_fc_LL:"
    it = _integrator()
    with pytest.raises(AssertionError, match="No interpolating function to save"):
        it.saveInterpolatingFn(str(tmp_path / "f.pkl"))
    it.bins, it.bin_midpoints = g[tag + 'edges'], g[tag + 'midpoints']
    it.integral, it.density = list(g[tag + 'integral']), g[tag + 'density']
    it.hasBins = it.hasMCIntegral = it.hasMCDensity = True
    it.makeInterpolatingFn(bounds=(it.bins[0], it.bins[-1]))
    it.makeInterpolatingDensity('lin')
    it.saveInterpolatingFn(str(tmp_path / "f.pkl"))
    it.saveInterpolatingDensity(str(tmp_path / "d.pkl"))
    other = _integrator()
    other.loadInterpolatingFn(str(tmp_path / "f.pkl"))
    other.loadInterpolatingDensity(str(tmp_path / "d.pkl"))
    assert other.hasInterpIntegral and other.hasInterpDensity
    x = np.linspace(it.bins[0], it.bins[-1], 57)
    assert_array_equal(other.interpFn(x), it.interpFn(x))
    assert_array_equal(other.interpDensity(x), it.interpDensity(x))
    assert other.interpDensity(it.bins[0] - 1.) == 0. and other.interpDensity(it.bins[-1] + 1.) == 0.

    # hooks: nothing by default; a subclass with an analytic integral gets its finite-difference density
    it.findAnalyticIntegral()
    assert it.hasAnalyticIntegral is False and it.analyticIntegral(.3) is None
    it.setAnalyticDensity('lin')
    assert it.hasAnalyticDensity is False

    class cubic(integrator):
        def findAnalyticIntegral(self):
            self.hasAnalyticIntegral = True

        def analyticIntegral(self, obs):
            return np.asarray(obs) ** 3

    c = cubic()
    c.bins = np.linspace(0., 1., 101)
    c.bin_midpoints = (c.bins[1:] + c.bins[:-1]) / 2.
    c.findAnalyticIntegral()
    c.setAnalyticDensity('lin')
    assert c.hasAnalyticDensity
    assert_allclose(c.analyticDensity(c.bin_midpoints[1:-1]), 3. * c.bin_midpoints[1:-1] ** 2, atol=2e-4)

    a = np.arange(24.).reshape(2, 3, 4)
    assert_array_equal(multidim_cumsum(a), a.cumsum(-1).cumsum(-2).cumsum(-3))
