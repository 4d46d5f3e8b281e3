"""samples_from_cdf (ref: utils/montecarlo_utils.py:93-375): the product's walk through the reference's list of cdf
shapes, against the vectors the reference itself produced with pynverse refusing every cdf.  The draw on the device
(inverse_transform_samples, checked on the GPU in test_gpu_callers.py) is replaced by the same interpolation on the
reference's np.random stream."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle import jmc_oracle as O
from test_oracle_golden import _cdf_cases


def test_every_branch_matches_the_reference(golden, monkeypatch):
    from jetmontecarlo_b200.utils import montecarlo_utils as MU
    g = golden("samples_from_cdf")
    draws = []

    def fake_draw(cdf_vals, x_vals, num_samples, seed=None):
        draws.append(len(cdf_vals))
        return O.inverse_transform(cdf_vals, x_vals, np.random.RandomState(seed).rand(num_samples))

    monkeypatch.setattr(MU, "inverse_transform_samples", fake_draw)
    for name, (cdf, domain, kw) in _cdf_cases().items():
        draws.clear()
        samples, weights = MU.samples_from_cdf(cdf, 1500, domain=domain, verbose=0, seed=int(g[name + ':seed']), **kw)
        assert_array_equal(samples, g[name + ':samples'], err_msg=name)
        assert_array_equal(weights, g[name + ':weights'], err_msg=name)
        assert len(draws) == (0 if name == 'starts_at_one' else 1), name     # zero bin: nothing is drawn
    for name in ('wiggle_forced', 'turnpoint'):
        cdf, domain, _ = _cdf_cases()[name]
        with pytest.raises(ValueError, match="not monotone"):
            MU.samples_from_cdf(cdf, 10, domain=domain, verbose=0, seed=1)
    with pytest.raises(AssertionError, match="domain"):
        MU.samples_from_cdf(lambda x: x, 10)
    # overflow bin: a cdf that never leaves 0 (with a dip so that it is not trivially monotone)
    tiny = lambda x: 1e-30 * (2. + np.sin(50 * x))
    s, w = MU.samples_from_cdf(tiny, 7, domain=[0, 2.], verbose=0, seed=1)
    assert_array_equal(s, np.full(7, 2.)) and np.all(w == 1)
    s_ref, _ = O.samples_from_cdf_tabulated(tiny, 7, [0, 2.], np.random.RandomState(1).rand)
    assert_array_equal(s, s_ref)


def test_probe_points_are_the_reference_grid(golden):
    from jetmontecarlo_b200.utils.montecarlo_utils import _cdf_probe_points
    for domain in ([0, 1], [1e-8, 1.], [-4., 4.], [0., .1]):
        assert_array_equal(_cdf_probe_points(domain), O.cdf_probe_points(domain))
    assert _cdf_probe_points([0, 1])[0] == 0 and len(_cdf_probe_points([0, 1])) == 1001
