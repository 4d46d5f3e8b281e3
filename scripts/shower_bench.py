import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gpu_helpers as H
from jetmontecarlo_b200.fused import Integrand
edges = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
for jet in ('quark', 'gluon'):
    for acc in ('LL', 'MLL'):
        for prec, n in (('f64', 20_000_000), ('f32', 50_000_000)):
            ig = Integrand.shower(beta=2., jet_type=jet, acc=acc, groomer='softdrop', z_cut=.1, obs_acc='MLL', n_emissions=2)
            H.run(ig, 1000000, 1, edges, prec)
            torch.cuda.synchronize(); t = time.perf_counter()
            r = H.run(ig, n, 2, edges, prec)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            print(f"shower {jet:5s} {acc:3s} {prec}: {n/dt:.3e} jets/s  ({dt*1e3:.1f} ms for {n:.0e}); in-range {int(r['count'].sum())}")
