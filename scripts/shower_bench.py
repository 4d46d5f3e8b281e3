"""The veto-shower kernels alone: jets/s of the FP32 kernel for the bench's three configurations (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gpu_helpers as H
from jetmontecarlo_b200.fused import Integrand
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5 * 10 ** 7
edges = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
cases = [("quark MLL ungroomed n=1", dict(jet_type='quark', acc='MLL', groomer='ungroomed', obs_acc='MLL', n_emissions=1)),
         ("gluon MLL ungroomed n=1", dict(jet_type='gluon', acc='MLL', groomer='ungroomed', obs_acc='MLL', n_emissions=1)),
         ("quark MLL softdrop n=2", dict(jet_type='quark', acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL', n_emissions=2)),
         ("quark MLL rss n=2", dict(jet_type='quark', acc='MLL', groomer='rss', z_cut=.1, f=1., obs_acc='MLL', n_emissions=2)),
         ("quark LL ungroomed n=1", dict(jet_type='quark', acc='LL', groomer='ungroomed', obs_acc='LL', n_emissions=1))]
for label, kw in cases:
    ig = Integrand.shower(beta=2., **kw)
    best = 1e9
    for k in range(4):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); out = H.run(ig, n, 10 + k, edges, 'f32'); ev[1].record(); torch.cuda.synchronize()
        if k: best = min(best, ev[0].elapsed_time(ev[1]))
    print(f"{label:28s} {best:8.3f} ms  {n / best * 1e3:.3e} jets/s   (binned %d)" % int(out['count'].sum()))
