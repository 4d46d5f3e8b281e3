"""The small device entry points the Python mirror uses instead of array-library arithmetic: fills, tile,
finite count, and the packing of the two collectives' buffers (checked here on one GPU against NumPy; the 2-rank
test exercises them around a real all-reduce)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fill_memset_tile_count_finite(cuda):
    import torch
    from jetmontecarlo_b200 import _device as D, _lib
    L = _lib.lib
    z = D.zeros((1000, 3))
    assert float(z.abs().sum()) == 0. and z.dtype == torch.float64
    zi = D.zeros((77,), dtype=torch.int64)
    assert int(zi.sum()) == 0
    f = D.full((12345,), np.nan)
    assert bool(torch.isnan(f).all())
    f = D.full((5, 7), -2.5)
    np.testing.assert_array_equal(D.to_host(f), np.full((5, 7), -2.5))
    a = np.random.RandomState(0).standard_normal(1001)
    out = D.empty((3003,))
    _lib.check(L.jmc_tile(D.ptr(D.to_device(a)), 1001, 3, D.ptr(out), D.stream()))
    np.testing.assert_array_equal(D.to_host(out), np.tile(a, 3))
    b = a.copy()
    b[::7] = np.nan
    b[3::11] = np.inf
    n_finite = C.c_uint64(0)
    _lib.check(L.jmc_count_finite(D.ptr(D.to_device(b)), len(b), C.byref(n_finite), D.stream()))
    assert n_finite.value == int(np.isfinite(b).sum())
    assert L.jmc_fill(None, 5, 1., D.stream()) < 0 and L.jmc_tile(None, 5, 0, None, D.stream()) < 0


def test_vals_to_pdf_without_finite_values(cuda):
    from jetmontecarlo_b200.utils.hist_utils import vals_to_pdf
    with pytest.warns(UserWarning, match="No finite values"):
        xs, pdf = vals_to_pdf(np.full(100, np.nan), 20, log_cutoff=1e-10)
    assert len(xs) == 0 and len(pdf) == 0


def test_collective_buffers(cuda):
    """[sum w | sum w^2 | count] and [-min | max | has-NaN] rows, there and back; the same functions on host tensors
    (the gloo form) give the same buffers"""
    import torch
    from jetmontecarlo_b200 import _device as D, distributed as du
    rng = np.random.RandomState(1)
    sw, sw2 = rng.standard_normal(99), rng.random(99)
    cnt = rng.randint(0, 2 ** 40, 99).astype(np.int64)
    d = [D.to_device(sw), D.to_device(sw2), torch.from_numpy(cnt).cuda()]
    buf = du.pack_histograms(*d)
    h = [torch.from_numpy(sw), torch.from_numpy(sw2), torch.from_numpy(cnt)]
    np.testing.assert_array_equal(buf.cpu().numpy(), du.pack_histograms(*h).numpy())
    buf *= 2                                        # "all-reduce" over two identical ranks
    du.unpack_histograms(buf, *d)
    np.testing.assert_array_equal(d[0].cpu().numpy(), 2 * sw)
    np.testing.assert_array_equal(d[1].cpu().numpy(), 2 * sw2)
    np.testing.assert_array_equal(d[2].cpu().numpy(), 2 * cnt)
    from jetmontecarlo_b200 import _lib
    mm = np.array([[1e-3, 2.], [np.nan, 5.], [np.inf, -np.inf], [-4., np.nan]])
    t = D.to_device(mm.copy())
    packed = D.empty((4, 3))
    _lib.check(_lib.lib.jmc_pack_minmax(D.ptr(t), 4, D.ptr(packed), 0, D.stream()))
    p = packed.cpu().numpy()
    np.testing.assert_array_equal(p[0], [-1e-3, 2., 0.])
    assert p[1][2] == 1. and p[1][0] == -np.inf and p[3][2] == 1. and p[3][1] == -np.inf
    np.testing.assert_array_equal(p[2], [-np.inf, -np.inf, 0.])
    _lib.check(_lib.lib.jmc_pack_minmax(D.ptr(t), 4, D.ptr(packed), 1, D.stream()))
    back = t.cpu().numpy()
    np.testing.assert_array_equal(back[0], mm[0])
    assert np.all(np.isnan(back[1])) and np.all(np.isnan(back[3]))
    np.testing.assert_array_equal(back[2], mm[2])
