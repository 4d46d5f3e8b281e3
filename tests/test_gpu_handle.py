"""GPU tests of the jmc_handle entry points (cached plan, device workspace, pinned staging), of the counts-on-request
mode of the FP32 kernel and of the pair layout of its sample streams (odd index ranges)."""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

pytestmark = pytest.mark.gpu

EDGES = np.logspace(-11, np.log10(.5), 100)


def _igs():
    from jetmontecarlo_b200.fused import Integrand
    return [Integrand.radiator('log', 1e-10, beta=2., jet_type='gluon', fixedcoupling=False, obs_acc='MLL', split_acc='MLL'),
            Integrand.critical('log', .1, 1e-10, fixedcoupling=False),
            Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False),
            Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=True),
            Integrand.pre_crit_sub('log', .1, 1e-10)]


def _handle_run(h, ig, n, seed, edges, precision, first=0, want_minmax=False):
    import torch
    from jetmontecarlo_b200 import _device as D, _lib
    s = ig.c_struct()
    nb = 0 if edges is None else len(edges) - 1
    sw, sw2, cnt = D.zeros((max(nb, 1),)), D.zeros((max(nb, 1),)), D.zeros((max(nb, 1),), dtype=torch.int64)
    mm = torch.tensor([np.inf, -np.inf], dtype=torch.float64, device="cuda") if want_minmax else None
    _lib.check(_lib.lib.jmc_handle_run(h.ptr, C.byref(s), n, seed, first, {'f64': 0, 'f32': 1}[precision],
                                       None if edges is None else edges.ctypes.data_as(C.c_void_p), nb + 1 if nb else 0,
                                       D.ptr(sw if nb else None), D.ptr(sw2 if nb else None), D.ptr(cnt if nb else None),
                                       D.ptr(mm), D.stream()))
    torch.cuda.synchronize()
    return dict(sum_w=sw.cpu().numpy(), sum_w2=sw2.cpu().numpy(), count=cnt.cpu().numpy(),
                minmax=None if mm is None else mm.cpu().numpy())


@pytest.mark.parametrize("precision", ['f32', 'f64'])
def test_handle_run_equals_jmc_run(cuda, precision):
    """host edges through the handle == device edges through jmc_run: same kernels, same streams"""
    import gpu_helpers as H
    from jetmontecarlo_b200 import _lib
    h = _lib.Handle()
    n = 300_001
    for ig in _igs():
        ref = H.run(ig, n, 12, EDGES, precision, first=7)
        for rep in range(2):                       # second call: the cached job
            got = _handle_run(h, ig, n, 12, EDGES, precision, first=7)
            assert_array_equal(got['count'], ref['count'].cpu().numpy())
            assert_allclose(got['sum_w'], ref['sum_w'].cpu().numpy(), rtol=2e-6, atol=1e-9 * np.abs(got['sum_w']).max())
            assert_allclose(got['sum_w2'], ref['sum_w2'].cpu().numpy(), rtol=4e-6, atol=1e-9 * np.abs(got['sum_w2']).max())
    # new edges under the same handle, then the min/max-only pass
    ig = _igs()[2]
    e2 = np.logspace(-8, -1, 41)
    got = _handle_run(h, ig, n, 3, e2, precision)
    assert_array_equal(got['count'], H.run(ig, n, 3, e2, precision)['count'].cpu().numpy())
    mm = _handle_run(h, ig, n, 3, None, precision, want_minmax=True)['minmax']
    assert_allclose(mm, H.run(ig, n, 3, None, precision, want_minmax=True)['minmax'], rtol=1e-12)
    h.close()


def test_handle_integrate_host(cuda):
    """whole job host -> host on an explicit handle: equals fused.integrate, call after call, also after the edges
    or the integrand change; the implicit-handle form jmc_integrate_host gives the same"""
    from jetmontecarlo_b200 import _lib, fused
    h = _lib.Handle()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for ig, bc, edges in ((_igs()[2], [1., 'minus'], EDGES), (_igs()[0], [0., 'plus'], EDGES),
                          (_igs()[0], [0., 'plus'], np.logspace(-9, np.log10(.5), 64)), (_igs()[3], [1., 'minus'], EDGES)):
        nb = len(edges) - 1
        for seed in (5, 6):
            out = [np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb), np.empty(nb), np.empty(nb),
                   np.empty(nb, dtype=np.uint64)]
            s = ig.c_struct()
            kind = _lib.BC_LAST_MINUS if bc[1] == 'minus' else _lib.BC_LAST_PLUS
            for entry in ("handle", "implicit"):
                if entry == "handle":
                    rc = _lib.lib.jmc_handle_integrate_host(h.ptr, C.byref(s), 200_000, seed, 0, _lib.F32, P(edges),
                                                            nb + 1, kind, bc[0], *[P(o) for o in out])
                else:
                    rc = _lib.lib.jmc_integrate_host(C.byref(s), 200_000, seed, 0, _lib.F32, P(edges), nb + 1, kind,
                                                     bc[0], *[P(o) for o in out])
                _lib.check(rc)
                it = fused.integrate(ig, 200_000, edges=edges, seed=seed, precision='f32', last_bc=bc)
                assert_array_equal(out[6].astype(np.int64), it._count.cpu().numpy())
                assert_allclose(out[0], it.density, rtol=5e-6, atol=1e-9 * np.abs(it.density).max())
                assert_allclose(out[2], np.asarray(it.integral), rtol=5e-6, atol=1e-8)
                assert_allclose(out[3], it.integralErr, rtol=5e-6, atol=1e-8)
    h.close()


@pytest.mark.parametrize("which", [1, 2])
def test_counts_on_request(cuda, which):
    """counts='weighted': the count histogram holds only the samples with a non-zero weight; sum w, sum w^2 and with
    them density, error and integral are those of the counted mode (the reference uses the count only to zero the
    density of bins without samples, integrator.py:101,123-124)"""
    import gpu_helpers as H
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    n, seed = 2_000_001, 21
    mk = (lambda c: Integrand.critical('log', .1, 1e-10, fixedcoupling=False, counts=c)) if which == 1 else \
         (lambda c: Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False, counts=c))
    a, w = H.run(mk('all'), n, seed, EDGES, 'f32', first=3), H.run(mk('weighted'), n, seed, EDGES, 'f32', first=3)
    assert_allclose(w['sum_w'].cpu().numpy(), a['sum_w'].cpu().numpy(), rtol=2e-6)
    assert_allclose(w['sum_w2'].cpu().numpy(), a['sum_w2'].cpu().numpy(), rtol=4e-6)
    ca, cw = a['count'].cpu().numpy(), w['count'].cpu().numpy()
    # (critical: the samples inside the binned range are mostly above the Landau scale, so the two counts are close)
    assert np.all(cw <= ca) and cw.sum() < (1. if which == 1 else .01) * ca.sum()
    _, obs, wt = H.dump(mk('all'), n, seed, 'f32', first=3)
    ref = np.histogram(obs[wt != 0], EDGES)[0]
    assert abs(cw - ref).sum() <= 3 + 2e-5 * ref.sum()          # FP32 edge ties (the dump is another instantiation)
    assert np.all((cw > 0) == (w['sum_w2'].cpu().numpy() > 0))
    ia = fused.integrate(mk('all'), n, edges=EDGES, seed=seed, precision='f32', last_bc=[1., 'minus'])
    iw = fused.integrate(mk('weighted'), n, edges=EDGES, seed=seed, precision='f32', last_bc=[1., 'minus'])
    assert_allclose(iw.density, ia.density, rtol=2e-6)
    assert_allclose(np.asarray(iw.integral), np.asarray(ia.integral), rtol=2e-6)
    assert_allclose(iw.integralErr, ia.integralErr, rtol=4e-6)


def test_odd_index_ranges(cuda):
    """the FP32 streams are laid out in pairs of samples (one Philox block per pair for the two-uniform kinds); a
    range that starts or ends inside a pair is the histogram of exactly its own samples"""
    import gpu_helpers as H
    for ig in _igs():
        _, obs_all, w_all = H.dump(ig, 4001, 9, 'f32', first=0)
        for first, n in ((1, 1), (1, 2), (0, 1), (3, 1000), (2, 999), (1, 3999), (0, 4001), (7, 3)):
            r = H.run(ig, n, 9, EDGES, 'f32', first=first)
            o, wt = obs_all[first:first + n], w_all[first:first + n]
            assert abs(r['count'].cpu().numpy() - np.histogram(o, EDGES)[0]).sum() <= 1, (ig.kind, first, n)
            ref = np.histogram(o, EDGES, weights=np.nan_to_num(wt))[0]
            assert_allclose(r['sum_w'].cpu().numpy(), ref, rtol=1e-4, atol=1e-6 * max(np.abs(ref).max(), 1e-300))
        # and a dump that starts inside a pair is the tail of the dump that starts at the pair
        _, obs_odd, _ = H.dump(ig, 100, 9, 'f32', first=1)
        assert_array_equal(obs_odd, obs_all[1:101])
