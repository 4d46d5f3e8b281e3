"""GPU tests of the generic MonteCarloProcess and the EWOC built on it.

Replay parity: the reference's own accepted samples (golden vectors made by running the unmodified reference) go
through the device kernels for the pairwise observables, the energy weights and the EWOC histogram and must reproduce
what the reference computed.  Fresh-sample parity: the device's hit-or-miss sampling has the reference's acceptance,
weights and distribution (same constraint, same cross section) within statistics, and the fused kernel equals the
staged path on its own stream.  ref: montecarlo/process.py:109-191, montecarlo/ewoc.py:11-75,
examples/processes/ee2hadrons_LO.py"""
import ctypes as C

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose, assert_array_equal

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


def _process_with_reference_samples(g, method):
    """an ee2hadrons_LO object that holds the reference's samples instead of its own"""
    from jetmontecarlo_b200 import _device as D
    from jetmontecarlo_b200.processes import ee2hadrons_LO as E
    p = E.ee2hadrons_LO(energy=1000., num_samples=16, sampleMethod=method, seed=1)
    p.samples = g['samples'].copy()
    p.jacobians = g['jacobians'].copy()
    p.device_samples = D.to_device(g['samples'])
    p.device_jacobians = D.to_device(g['jacobians'])
    p.num_rejected_samples = int(g['num_rejected'])
    p.setArea()
    return p, E


@pytest.mark.parametrize("method", ['log', 'lin'])
def test_pairs_and_ewoc_on_the_reference_samples(cuda, golden, method):
    g = golden(f"process_ee2h_{method}")
    p, E = _process_with_reference_samples(g, method)
    from jetmontecarlo_b200.montecarlo.ewoc import Npoint_MonteCarloEWOC
    assert p.area == float(g['area'])
    named = p.named_samples()
    for k in ('x_1', 'x_2', 'x_3'):
        assert_array_equal(np.array(named[k]), g[k])
    fns = dict(costheta=E.costheta_ij, m2=E.m2_ij, theta=E.theta_ij, z_i=E.z_i, z_j=E.z_j)
    obs = {}
    for name, fn in fns.items():
        ob = E.ee2hLO_PairwiseObservable(p, fn, name)
        assert ob.device_observables is not None                      # the kernel ran
        # arccos: CUDA's acos and libm's agree to 1 ulp, not to the bit
        assert_allclose(ob.observables, g[f'obs:{name}'], rtol=0 if name != 'theta' else 5e-16, atol=0, err_msg=name)
        assert_array_equal(ob.weights, g[f'obs:{name}:weights'])
        # and the functions themselves on host arrays (user-facing)
        x_1, x_2 = 1 - g['samples'][:, 0], 1 - g['samples'][:, 1]
        assert_array_equal(fn(x_1, x_2), g[f'obs:{name}'][:len(x_1)])
        obs[name] = ob
    ob = E.ee2hLO_PairwiseObservable(p, E.gecf(1.5), 'gecf')
    assert_allclose(ob.observables, g['obs:gecf1.5'], rtol=4e-16, atol=0)         # pow(): 1 ulp
    for name, spacing, ew in (('costheta', 'lin', (1, 1)), ('m2', 'lin', (1, 1)), ('m2', 'log', (2, 1)),
                              ('theta', 'log', (1, 1))):
        tag = f'ewoc:{name}:{spacing}:{ew[0]}{ew[1]}'
        ewoc = Npoint_MonteCarloEWOC(p, obs[name], energy_weights=ew)
        with pytest.raises(ValueError, match="Weights have not been set"):
            ewoc.generate_ewoc(numbins=40, binspacing=spacing)
        ewoc.set_weights(obs['z_i'], obs['z_j'])
        assert_array_equal(ewoc.weights, g[tag + ':weights'])
        ewoc.generate_ewoc(numbins=40, binspacing=spacing)
        tol = dict(rtol=1e-12, atol=0) if name == 'theta' else dict(rtol=0, atol=0)
        assert_allclose(ewoc.bins, g[tag + ':bins'], **tol)
        scale = np.abs(g[tag + ':density']).max()
        assert_allclose(ewoc.density, g[tag + ':density'], rtol=1e-12, atol=1e-13 * scale)
        # NumPy forms a weighted histogram as differences of a running cumsum: a bin whose sum of w^2 is 1e-8 of the
        # total carries 1e-16 / 1e-8 of relative noise in the reference's own value (the device sums each bin directly)
        assert_allclose(ewoc.densityErr, g[tag + ':density_err'], rtol=1e-6, atol=1e-13 * scale)
    ewoc = Npoint_MonteCarloEWOC(p, obs['costheta'])
    ewoc.set_weights(obs['z_i'], obs['z_j'])
    ewoc.generate_ewoc(numbins=25, binspacing='lin', minval=-1., maxval=1.)
    assert_array_equal(ewoc.bins, g['ewoc:range:bins'])
    assert_allclose(ewoc.density, g['ewoc:range:density'], rtol=1e-12, atol=1e-13 * np.abs(g['ewoc:range:density']).max())


@pytest.mark.parametrize("method", ['log', 'lin'])
def test_device_sampling_is_the_reference_algorithm(cuda, method):
    """jmc_process_sample on the device's own Philox stream == the oracle's hit-or-miss loop fed with the same
    uniforms stream by stream (sample i = first accepted trial of stream i): samples, weights, rejected count"""
    import gpu_helpers as H
    from jetmontecarlo_b200.processes import ee2hadrons_LO as E
    n, seed = 4000, 99
    p = E.ee2hadrons_LO(energy=1000., num_samples=n, sampleMethod=method, seed=seed)
    pre = O.ee2hadrons_lo_prefactor(1000.)
    bounds = [(0, 1), (0, 1)]
    u = H.uniforms(n, seed, 0, 64)                     # 32 trials of every stream
    ref_s, ref_w, rejected = [], [], 0
    for i in range(n):
        s, w, r = O.process_hit_or_miss(u[i], 1, method, 1e-8, bounds, O.ee2h_in_phasespace,
                                        lambda a, b: O.ee2h_differential_xsec(a, b, pre))
        ref_s.append(s[0]); ref_w.append(w[0]); rejected += r
    assert_allclose(p.samples, np.array(ref_s), rtol=4e-16, atol=0)            # exp(): 1 ulp
    assert_allclose(p.jacobians, np.array(ref_w), rtol=1e-14, atol=0)      # exp() twice, a product, a quotient
    assert p.num_rejected_samples == rejected
    assert p.area == O.process_area(bounds, n, rejected)
    assert (rejected > .8 * n) == (method == 'lin')                             # half of the linear box is outside
    # a second call is a new stream and appends nothing silently: the reference's list semantics
    with pytest.raises(AttributeError):
        p.generateSamples(10)


def test_fused_ewoc_equals_the_staged_path(cuda):
    """one kernel (generate -> pairs -> weights -> histogram) == sampling, pair kernels, weight kernel, jmc_hist on
    the same stream; and the multi-GPU sharding rule: index ranges add up"""
    from jetmontecarlo_b200 import _device as D, _lib
    from jetmontecarlo_b200.montecarlo.ewoc import Npoint_MonteCarloEWOC, fused_ewoc
    from jetmontecarlo_b200.processes import ee2hadrons_LO as E
    n, seed = 300_000, 5
    for method, fn, spacing, ew in (('log', E.costheta_ij, 'lin', (1, 1)), ('lin', E.m2_ij, 'lin', (2, 1)),
                                    ('log', E.theta_ij, 'log', (1, 1)), ('lin', E.gecf(1.5), 'log', (1, 2))):
        p = E.ee2hadrons_LO(energy=1000., num_samples=n, sampleMethod=method, seed=seed)
        ob = E.ee2hLO_PairwiseObservable(p, fn, 'o')
        ewoc = Npoint_MonteCarloEWOC(p, ob, energy_weights=ew)
        ewoc.set_weights(E.ee2hLO_PairwiseObservable(p, E.z_i, 'zi'), E.ee2hLO_PairwiseObservable(p, E.z_j, 'zj'))
        ewoc.generate_ewoc(numbins=60, binspacing=spacing)
        it = fused_ewoc(p, n, fn, energy_weights=ew, numbins=60, binspacing=spacing, seed=seed)
        assert it.num_rejected == p.num_rejected_samples and it.area == p.area
        assert_array_equal(it.bins, ewoc.bins)
        assert_array_equal(it._count.cpu().numpy(), ewoc._count.cpu().numpy())
        top = np.abs(ewoc.density).max()
        assert_allclose(it.density, ewoc.density, rtol=1e-11, atol=1e-13 * top)
        assert_allclose(it.densityErr, ewoc.densityErr, rtol=1e-11, atol=1e-13 * top)
    # additivity over index ranges (what sharding over GPUs relies on)
    p = E.ee2hadrons_LO(energy=1000., num_samples=8, sampleMethod='log', seed=1)
    desc = p._descriptor()
    edges = D.to_device(np.linspace(-1., 1., 41))

    def run(first, count):
        sw, sw2 = D.zeros((40,)), D.zeros((40,))
        cnt, rej = D.zeros((40,), dtype=torch.int64), D.zeros((1,), dtype=torch.int64)
        _lib.check(_lib.lib.jmc_process_ewoc(C.byref(desc), count, 7, first, _lib.PAIR_COSTHETA, 0., 1., 1., D.ptr(edges),
                                             41, D.ptr(sw), D.ptr(sw2), D.ptr(cnt), None, D.ptr(rej), D.stream()))
        return sw.cpu().numpy(), cnt.cpu().numpy(), int(rej.item())
    whole, a, b = run(0, 100_001), run(0, 33_333), run(33_333, 66_668)
    assert_array_equal(whole[1], a[1] + b[1])
    assert whole[2] == a[2] + b[2]
    assert_allclose(whole[0], a[0] + b[0], rtol=1e-12)


def test_process_statistics_against_the_reference_run(cuda, golden):
    """fresh device samples against the reference's own run (3000 samples): the acceptance agrees within the
    reference's binomial error, and the EWOC agrees with a CPU run of the oracle's loop (2e5 samples, NumPy uniforms)
    within the combined statistical errors"""
    from jetmontecarlo_b200.montecarlo.ewoc import fused_ewoc
    from jetmontecarlo_b200.processes import ee2hadrons_LO as E
    pre = O.ee2hadrons_lo_prefactor(1000.)
    bounds = [(0, 1), (0, 1)]
    for method in ('log', 'lin'):
        g = golden(f"process_ee2h_{method}")
        p = E.ee2hadrons_LO(energy=1000., num_samples=8, sampleMethod=method, seed=3)
        edges = np.linspace(0., 1., 21)
        it = fused_ewoc(p, 4_000_000, E.m2_ij, binspacing='lin', edges=edges, seed=11)
        eff_ref = 3000. / (3000. + float(g['num_rejected']))
        eff = 4e6 / (4e6 + it.num_rejected)
        assert abs(eff - eff_ref) < 5 * np.sqrt(eff_ref * (1 - eff_ref) / 3000.) + 1e-3
        if method == 'lin':
            continue          # 1 / ((1 - x_1)(1 - x_2)) weights on uniform points: no usable error estimate at 2e5 samples
        n_cpu = 200_000
        u = np.random.RandomState(5).rand(2 * n_cpu + 4000)
        s, w, rej = O.process_hit_or_miss(u, n_cpu, method, 1e-8, bounds, O.ee2h_in_phasespace,
                                          lambda a, b: O.ee2h_differential_xsec(a, b, pre))
        obs = O.ee2h_pair_observable(s, 'm2')
        wts = O.ewoc_weights(O.ee2h_pair_observable(s, 'z_i'), O.ee2h_pair_observable(s, 'z_j'), np.tile(w, 3))
        h, h2, c = O.histograms(obs, wts, edges)
        dens, err = O.density_from_hist(h, h2, c, edges, O.process_area(bounds, n_cpu, rej), len(obs))
        ok = (c > 200) & (err > 0)
        pull = (it.density[ok] - dens[ok]) / np.sqrt(err[ok] ** 2 + it.densityErr[ok] ** 2)
        assert ok.sum() >= 10 and np.median(np.abs(pull)) < 1.5 and np.abs(pull).max() < 6., np.abs(pull)


def test_process_api_and_errors(cuda):
    from jetmontecarlo_b200.montecarlo.process import BasicProcess, MonteCarloObservable, MonteCarloProcess
    from jetmontecarlo_b200.processes import ee2hadrons_LO as E
    assert str(BasicProcess("test", 1.0)) == "The process [test] at an energy of 1.0 GeV."
    p = E.ee2hadrons_LO(energy=1000., num_samples=100, seed=2)
    assert p.kinematic_vars == ['1 - x_1', '1 - x_2'] and p.num_variables == 2 and p.accuracy == 'LO'
    assert p.samples.shape == (100, 2) and len(p.jacobians) == 100 and 0 < p.area <= 1.
    assert p.lies_in_phasespace(**{'1 - x_1': .2, '1 - x_2': .3}) and not p.lies_in_phasespace(**{'1 - x_1': .7, '1 - x_2': .6})
    with pytest.raises(AssertionError, match="not"):
        p.lies_in_phasespace(x=1.)
    # the Python callbacks give what the device gave
    a, b = p.samples[0]
    jac = a * b
    assert_allclose(jac * p.differential_xsec(**{'1 - x_1': a, '1 - x_2': b}), p.jacobians[0], rtol=1e-14)
    named = p.named_samples()
    assert_allclose(np.array(named['x_1']) + np.array(named['x_2']) + np.array(named['x_3']), 2., rtol=1e-15)
    generic = MonteCarloObservable(p, lambda a, b: a + b, 'sum')
    assert_allclose(generic.observables, p.samples.sum(axis=1)) and generic.weights is p.jacobians

    class NoDevice(MonteCarloProcess):
        def _phasespace_constraints(self, **kw):
            return True

        def _differential_xsec(self, **kw):
            return 1.
    q = NoDevice(kinematic_bounds=[(0, 1)], sampleMethod='lin', kinematic_vars=['x'], name='n', energy=1.)
    assert q.area is None
    with pytest.raises(NotImplementedError, match="no device implementation"):
        q.generateSamples(10)
    with pytest.raises(AssertionError, match="number of bounds"):
        NoDevice(kinematic_bounds=[(0, 1), (0, 1)], sampleMethod='lin', kinematic_vars=['x'], name='n', energy=1.)
