"""Where the wall time of BASELINE config 5 goes (5000 theta_crit points x 5e6 samples x 5000 bins): stage timers
around the steps of fused.integrate_batch, device work synchronised at every stage."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jetmontecarlo_b200 import fused, _lib, _device as D
from jetmontecarlo_b200.fused import Integrand, _as_struct_array
from jetmontecarlo_b200.utils.interpolation import lin_log_mixed_list
npts, nbins, n5 = 5000, 5000, 5_000_000
grid = lin_log_mixed_list(1e-10, 1., npts)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.) + time.perf_counter() - t0; return time.perf_counter()
t = time.perf_counter()
igs = [Integrand.radiator('log', 1e-8, beta=2., jet_type='quark', fixedcoupling=True, theta_scale=float(x)) for x in grid]
t = tick("build Integrand objects", t)
fused.integrate_batch(igs[:4], 100000, numbins=nbins, binspacing='log', min_log_bins=grid[:4] ** 2. * 1e-15, seed=1, precision='f32', last_bc=[0., 'plus'])
for rep in range(2):
    T.clear(); t = time.perf_counter(); t_all = t
    arr = _as_struct_array(igs); t = tick("struct array", t)
    mm = torch.empty((npts, 2), dtype=torch.float64, device=D.device()); mm[:, 0], mm[:, 1] = float('inf'), float('-inf')
    _lib.check(_lib.lib.jmc_run_batch(arr, npts, n5, 1, 0, _lib.F32, None, 0, None, None, None, D.ptr(mm), D.stream()))
    t = tick("min/max pass (jmc_run_batch)", t)
    mmh = D.to_host(mm); mlb = grid ** 2. * 1e-15
    edges = fused._set_bins_rows(mmh[:, 0], mmh[:, 1], mlb, nbins, 'log')
    t = tick("setBins rows on the host", t)
    ne = nbins; nb = ne - 1
    dens, err, integ, ierr = D.empty((npts, nb)), D.empty((npts, nb)), D.empty((npts, ne)), D.empty((npts, nb))
    count = D.zeros((npts, nb), dtype=torch.int64); sum_w, sum_w2 = D.zeros((npts, nb)), D.zeros((npts, nb))
    edges_dev = D.to_device(edges)
    t = tick("allocate + upload edges", t)
    _lib.check(_lib.lib.jmc_run_batch(arr, npts, n5, 1, 0, _lib.F32, edges.ctypes.data_as(C.c_void_p), ne,
                                      D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), None, D.stream()))
    t = tick("histogram pass (jmc_run_batch)", t)
    areas = D.to_device(np.array([ig.area for ig in igs], dtype=np.float64))
    _lib.check(_lib.lib.jmc_finalize_batch(npts, D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), D.ptr(edges_dev), ne,
                                           D.ptr(areas), n5, _lib.BC_LAST_PLUS, 0., D.ptr(dens), D.ptr(err), D.ptr(integ), D.ptr(ierr), D.stream()))
    t = tick("finalize", t)
    out = {k: D.to_host(v) for k, v in dict(density=dens, densityErr=err, integral=integ, integralErr=ierr, count=count).items()}
    t = tick("results to the host (5 arrays, 1 GB)", t)
    total = time.perf_counter() - t_all
print("total %.3f s" % total)
for k, v in T.items(): print("  %-34s %8.1f ms" % (k, v * 1e3))
