"""hist_kernel / minmax_kernel alone (CUDA events), arrays larger than L2: achieved HBM bandwidth vs the measured peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from jetmontecarlo_b200 import _device as D, _lib
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
g = torch.Generator(device="cuda").manual_seed(1)
obs = torch.exp(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * np.log(1e-11) )      # log-uniform in (1e-11, 1)
for label, w in (("all weights non-zero", torch.rand(n, dtype=torch.float64, device="cuda", generator=g)),
                 ("99.7 % zero weights (C3 rc)", None)):
    if w is None:
        w = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)
        w[torch.rand(n, device="cuda", generator=g) < .997] = 0.
    for ne, edges in ((100, np.logspace(-11, np.log10(.5), 100)), (5000, np.logspace(-11, np.log10(.5), 5000)),
                      (100, np.sort(np.exp(np.random.RandomState(0).uniform(np.log(1e-11), 0, 100)))),
                      (100, np.linspace(0., 1., 100)),
                      (101, np.insert(np.logspace(-10, 0, 100), 0, 1e-100))):
        e = D.to_device(edges); nb = len(edges) - 1
        lin = bool(np.allclose(np.diff(edges), np.diff(edges)[0]))
        if lin and 'obs_lin' not in globals():
            obs_lin = torch.rand(n, dtype=torch.float64, device="cuda", generator=g)     # uniform in (0, 1) for linear bins
        data = obs_lin if lin else obs
        sw, sw2, cnt = D.zeros((nb,)), D.zeros((nb,)), D.zeros((nb,), dtype=torch.int64)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(3):
            ev[0].record()
            _lib.check(_lib.lib.jmc_hist(D.ptr(data), D.ptr(w), n, D.ptr(e), len(edges), D.ptr(sw), D.ptr(sw2), D.ptr(cnt), D.stream()))
            ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        with np.errstate(all='ignore'):
            uni = ("uniform-lin" if np.allclose(np.diff(edges), np.diff(edges)[0]) else
                   "uniform-log" if np.allclose(np.diff(np.log(edges)), np.diff(np.log(edges))[0]) else
                   "zerobin+log" if np.allclose(np.diff(np.log(edges[1:])), np.diff(np.log(edges[1:]))[0]) else "irregular")
        print(f"jmc_hist  {label:28s} {ne:5d} {uni:11s} edges: {ms:7.3f} ms  {16*n/ms/1e6:7.0f} GB/s = {16*n/ms/1e6/peak:.2f} of measured HBM peak  ({n/ms*1e3:.3e} samples/s)")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(3):
    ev[0].record(); tot = obs.sum(); ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]); print(f"yardstick: torch sum of the same array {ms:7.3f} ms  {8*n/ms/1e6:7.0f} GB/s = {8*n/ms/1e6/peak:.2f} of measured HBM peak")
mm = D.empty((2,))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(3):
    ev[0].record(); _lib.check(_lib.lib.jmc_minmax(D.ptr(obs), n, D.ptr(mm), D.stream())); ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]); print(f"jmc_minmax {ms:7.3f} ms  {8*n/ms/1e6:7.0f} GB/s = {8*n/ms/1e6/peak:.2f} of measured HBM peak")
