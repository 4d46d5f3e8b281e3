"""Host <-> device plumbing of the Python mirror: large results come back through pinned staging buffers in chunks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_large_results_come_back_intact(cuda):
    import torch
    from jetmontecarlo_b200 import _device as D
    g = torch.Generator(device="cuda").manual_seed(5)
    for dtype, shape in ((torch.float64, (3, 4_100_003)), (torch.int64, (9_000_001,)), (torch.float64, (10, 7))):
        t = (torch.rand(shape, dtype=torch.float64, device="cuda", generator=g) * 1e6).to(dtype)
        got = D.to_host(t)
        assert got.shape == tuple(shape) and got.dtype == t.cpu().numpy().dtype
        np.testing.assert_array_equal(got, t.cpu().numpy())
    # a non-contiguous view
    t = torch.rand((4096, 4096), dtype=torch.float64, device="cuda", generator=g).t()
    np.testing.assert_array_equal(D.to_host(t), t.cpu().numpy())


def test_large_host_arrays_go_up_intact(cuda):
    import torch
    from jetmontecarlo_b200 import _device as D
    rng = np.random.RandomState(3)
    for shape in ((9_000_001,), (4_100_003, 2), (1000,)):
        a = rng.standard_normal(shape)
        t = D.to_device(a)
        assert t.is_cuda and t.dtype == torch.float64 and tuple(t.shape) == tuple(shape)
        np.testing.assert_array_equal(t.cpu().numpy(), a)
    # a strided host view (a column of an (N, 2) sample array) and a following kernel on torch's stream
    a = rng.standard_normal((9_000_001, 2))
    col = D.to_device(a[:, 1])
    np.testing.assert_array_equal(D.to_host(col), a[:, 1])
    from jetmontecarlo_b200._arrayfn import eval_fn
    from jetmontecarlo_b200 import _lib
    np.testing.assert_array_equal(np.asarray(eval_fn(_lib.FN_MUL, [a[:, 0], 2.])), a[:, 0] * 2.)
