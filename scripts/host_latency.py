"""Where the wall time of a small job through the host-buffer entry point goes (C1, 1e6 samples, 99 bins):
the whole call, and its stages issued separately with a synchronisation after each."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jetmontecarlo_b200 import _device as D, _lib
from jetmontecarlo_b200.fused import Integrand
L = _lib.lib
e1 = np.logspace(-11, np.log10(.5), 100); nb = 99
P = lambda a: a.ctypes.data_as(C.c_void_p)
bufs = [np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb), np.empty(nb), np.empty(nb), np.empty(nb, dtype=np.uint64)]
def timeit(fn, reps=300, warm=30):
    for k in range(warm): fn(k)
    torch.cuda.synchronize(); t = time.perf_counter()
    for k in range(reps): fn(warm + k)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e6
for name, ig, bc in (("C1 radiator fc (nan_to_num off: poison pass)", Integrand.radiator('log', 1e-10, beta=2.), _lib.BC_LAST_PLUS),
                     ("C1 radiator fc, nan_to_num on", Integrand.radiator('log', 1e-10, beta=2., nan_to_num=True), _lib.BC_LAST_PLUS),
                     ("C3 crit x sub rc, counts on request", Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False, counts='weighted'), _lib.BC_LAST_MINUS)):
    s = ig.c_struct()
    for n in (10**4, 10**6, 10**7):
        whole = timeit(lambda k: _lib.check(L.jmc_integrate_host(C.byref(s), n, 1, k * n, _lib.F32, P(e1), nb + 1, bc, 0., *[P(b) for b in bufs])))
        h = _lib.default_handle(0)
        sw, sw2, cnt = D.zeros((nb,)), D.zeros((nb,)), D.zeros((nb,), dtype=torch.int64)
        out = [D.empty((nb,)), D.empty((nb,)), D.empty((nb + 1,)), D.empty((nb,))]
        ed = D.to_device(e1)
        st = D.stream()
        def run(k):
            _lib.check(L.jmc_handle_run(h.ptr, C.byref(s), n, 1, k * n, _lib.F32, P(e1), nb + 1, D.ptr(sw), D.ptr(sw2), D.ptr(cnt), None, st))
            torch.cuda.synchronize()
        def fin(k):
            _lib.check(L.jmc_finalize(D.ptr(sw), D.ptr(sw2), D.ptr(cnt), D.ptr(ed), nb + 1, 1., n, bc, 0., *[D.ptr(o) for o in out], st))
            torch.cuda.synchronize()
        def sync_only(k):
            torch.cuda.synchronize()
        def memset(k):
            sw.zero_(); torch.cuda.synchronize()
        t_run, t_fin, t_sync, t_ms = timeit(run), timeit(fin), timeit(sync_only), timeit(memset)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for k in range(200):
            _lib.check(L.jmc_handle_run(h.ptr, C.byref(s), n, 1, k * n, _lib.F32, P(e1), nb + 1, D.ptr(sw), D.ptr(sw2), D.ptr(cnt), None, st))
        ev1.record(); torch.cuda.synchronize()
        # device time of ONE run (events enqueued right around it; the stream is idle before)
        dts = []
        for k in range(30):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(L.jmc_handle_run(h.ptr, C.byref(s), n, 1, k * n, _lib.F32, P(e1), nb + 1, D.ptr(sw), D.ptr(sw2), D.ptr(cnt), None, st))
            b.record(); torch.cuda.synchronize()
            dts.append(a.elapsed_time(b) * 1e3)
        dts.sort()
        print(f"    device time of one run (event to event, median of 30): {dts[15]:.1f} us  (spread shift env: {os.environ.get('JMC_FAST_SPREAD_SHIFT', 'default')})")
        print(f"{name:45s} n={n:.0e}: jmc_integrate_host {whole:7.1f} us | run+sync {t_run:6.1f}  finalize+sync {t_fin:6.1f}  "
              f"torch memset+sync {t_ms:6.1f}  bare sync {t_sync:5.1f} | run back-to-back (GPU time per job) {ev0.elapsed_time(ev1) * 5:6.1f} us")
