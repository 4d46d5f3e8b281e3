"""Fused FP32 kernel against the number of bins (block histogram size, lane-interleaved copies): C1 radiator, 1e9
samples, single job; and the batched grid kernel at C5's shape."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gpu_helpers as H
from jetmontecarlo_b200 import _lib, _device as D
from jetmontecarlo_b200.fused import Integrand, _as_struct_array
ig = Integrand.radiator('log', 1e-8, beta=2., jet_type='quark', fixedcoupling=True)
n = 10 ** 9
for ne in (100, 400, 1000, 2000, 5000):
    edges = np.logspace(-17, 0, ne)
    best = 1e9
    for k in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); H.run(ig, n, 5 + k, edges, 'f32'); ev[1].record(); torch.cuda.synchronize()
        if k: best = min(best, ev[0].elapsed_time(ev[1]))
    print(f"single job, {ne:5d} edges: {best:8.2f} ms  {n / best * 1e3:.3e} samples/s")
for npts, n5, ne in ((5000, 5_000_000, 5000), (5000, 5_000_000, 100), (500, 50_000_000, 5000)):
    grid = np.logspace(-10, 0, npts)
    igs = [Integrand.radiator('log', 1e-8, beta=2., jet_type='quark', fixedcoupling=True, theta_scale=float(x)) for x in grid]
    arr = _as_struct_array(igs)
    edges = np.array([np.logspace(-17, 0, ne) * x ** 2 for x in grid])
    nb = ne - 1
    count = D.zeros((npts, nb), dtype=torch.int64); sum_w, sum_w2 = D.zeros((npts, nb)), D.zeros((npts, nb))
    for k in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.check(_lib.lib.jmc_run_batch(arr, npts, n5, 1, 0, _lib.F32, edges.ctypes.data_as(C.c_void_p), ne,
                                          D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), None, D.stream()))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"batch {npts} points x {n5:.0e} samples x {ne} edges: {dt * 1e3:8.1f} ms  {npts * n5 / dt:.3e} samples/s")
