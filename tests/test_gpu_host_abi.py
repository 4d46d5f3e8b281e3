"""The host-buffer entry points of the C ABI (what a reference-side ctypes binding calls):
plain NumPy buffers in and out, no torch involved in the calls."""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import jmc_oracle as O

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_set_density_host(cuda, golden):
    from jetmontecarlo_b200 import _lib
    g = golden("c3")
    tag = "quark_rc_MLL:"
    obs = np.ascontiguousarray(g[tag + 'obs'])
    wj = np.nan_to_num(g[tag + 'weights']) * g['jac_crit'] * g['jac_sub']
    edges = np.ascontiguousarray(g['edges_fixed'])
    nb = len(edges) - 1
    dens, err, integ, ierr = np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb)
    sw, sw2, cnt = np.empty(nb), np.empty(nb), np.empty(nb, dtype=np.uint64)
    area = float(g['area_crit']) * float(g['area_sub'])
    rc = _lib.lib.jmc_set_density_host(P(obs), P(wj), len(obs), P(edges), len(edges), area, _lib.BC_LAST_MINUS, 1.,
                                       P(dens), P(err), P(integ), P(ierr), P(sw), P(sw2), P(cnt))
    assert rc == 0, _lib.lib.jmc_last_error()
    ref = O.integrate_hist(obs, wj, edges, area, last_bc=[1., 'minus'])
    assert_array_equal(cnt.astype(np.int64), ref['count'])
    top = np.abs(wj).sum()
    assert_allclose(sw, ref['sum_w'], rtol=1e-12, atol=1e-13 * top)
    assert_allclose(dens, g[tag + 'density'], rtol=1e-12, atol=1e-13 * top * area / len(obs) / np.diff(edges).min())
    assert_allclose(integ, g[tag + 'integral'], rtol=1e-11, atol=1e-12)
    assert_allclose(ierr, g[tag + 'integral_err'], rtol=1e-7, atol=1e-9 * ierr.max())   # sqrt of a cancelling cumsum
    # optional outputs may be NULL; bad arguments come back as error codes with a message, never as exceptions
    rc = _lib.lib.jmc_set_density_host(P(obs), P(wj), len(obs), P(edges), len(edges), area, _lib.BC_LAST_MINUS, 1.,
                                       P(dens), P(err), P(integ), P(ierr), None, None, None)
    assert rc == 0
    rc = _lib.lib.jmc_set_density_host(P(obs), P(wj), len(obs), P(edges), 1, area, 0, 0., P(dens), P(err), P(integ),
                                       P(ierr), None, None, None)
    assert rc == -1 and b"Need bins" in _lib.lib.jmc_last_error()
    big = np.logspace(-12, 0, 20000)
    rc = _lib.lib.jmc_set_density_host(P(obs), P(wj), len(obs), P(big), len(big), area, 0, 0., None, None, None, None,
                                       None, None, None)
    assert rc == -4 and b"shared memory" in _lib.lib.jmc_last_error()          # JMC_ELIMIT, documented


def test_sample_and_eval_fn_host(cuda):
    from jetmontecarlo_b200 import _lib
    n, seed = 5000, 321
    s = _lib.Sampler(kind=_lib.SAMPLER_CRITICAL, method=_lib.LOG, epsilon=1e-8, zc=.1, radius=1.)
    samples, jac = np.empty((n, 2)), np.empty(n)
    assert _lib.lib.jmc_sample_host(C.byref(s), n, seed, 0, P(samples), P(jac)) == 0
    ref, jref, _ = O.sample_critical('log', O.philox_uniforms53(seed, 0, n, 2), .1, 1e-8)
    assert_allclose(samples, ref, rtol=1e-14)
    assert_array_equal(jac, (samples[:, 0] - .1) * samples[:, 1])
    # array function on strided host views of the (N,2) sample array, as the reference stores samples
    p = _lib.FnParams(jet=_lib.GLUON, acc=_lib.ACC_MLL, fixed_coupling=0, z_cut=.1, beta=2.)
    out = np.empty(n)
    z_ptr = C.c_void_p(samples.ctypes.data)
    th_ptr = C.c_void_p(samples.ctypes.data + 8)
    rc = _lib.lib.jmc_eval_fn_host(_lib.FN_CRIT_PDF_RC, C.byref(p), n, z_ptr, 2, th_ptr, 2, None, 1, None, 1, P(out))
    assert rc == 0, _lib.lib.jmc_last_error()
    with np.errstate(all='ignore'):
        want = O.pdf_crit_rc(samples[:, 0], samples[:, 1], .1, 'gluon')
    assert_array_equal(np.isnan(out), np.isnan(want))
    assert_allclose(out, want, rtol=1e-10, equal_nan=True)
    assert np.isnan(want).mean() > .3


def test_integrate_host(cuda):
    """the whole job through one C call with host buffers == the device-pointer path"""
    from jetmontecarlo_b200 import _lib, fused
    from jetmontecarlo_b200.fused import Integrand
    ig = Integrand.crit_sub('log', .1, 1e-10, jet_type='quark', fixedcoupling=True, obs_acc='MLL')
    s = ig.c_struct()
    edges = np.logspace(-11, np.log10(.5), 100)
    nb, n = 99, 3_000_000
    dens, err, integ, ierr = np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb)
    sw, sw2, cnt = np.empty(nb), np.empty(nb), np.empty(nb, dtype=np.uint64)
    for prec in ('f64', 'f32'):
        rc = _lib.lib.jmc_integrate_host(C.byref(s), n, 7, 0, fused.PRECISIONS[prec], P(edges), 100,
                                         _lib.BC_LAST_MINUS, 1., P(dens), P(err), P(integ), P(ierr), P(sw), P(sw2),
                                         P(cnt))
        assert rc == 0, _lib.lib.jmc_last_error()
        it = fused.integrate(ig, n, edges=edges, seed=7, precision=prec, last_bc=[1., 'minus'])
        assert_array_equal(cnt.astype(np.int64), it._count.cpu().numpy())
        assert_allclose(dens, it.density, rtol=1e-9, atol=1e-12 * np.abs(it.density).max())
        assert_allclose(integ, np.asarray(it.integral), rtol=1e-9, atol=1e-12)
        assert integ[-1] == 1. and np.all(np.diff(integ) >= 0)
