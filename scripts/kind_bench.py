"""Throughput of every fused integrand (FP32 and FP64 modes), 1e9 / 1e8 samples, 99 log bins."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import gpu_helpers as H
from jetmontecarlo_b200.fused import Integrand
e = np.logspace(-11, np.log10(.5), 100)
cases = [
 ("C1  radiator fc LL quark", Integrand.radiator('log', 1e-10, beta=2., jet_type='quark', fixedcoupling=True)),
 ("C2i radiator rc MLL gluon", Integrand.radiator('log', 1e-10, beta=2., jet_type='gluon', fixedcoupling=False, obs_acc='MLL', split_acc='MLL')),
 ("    critical fc", Integrand.critical('log', .1, 1e-10, fixedcoupling=True)),
 ("    critical rc (65% NaN->0)", Integrand.critical('log', .1, 1e-10, fixedcoupling=False)),
 ("C3  crit x sub rc MLL (99.7% NaN->0)", Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False)),
 ("C3  same, counts on request", Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False, counts='weighted')),
 ("    critical rc, counts on request", Integrand.critical('log', .1, 1e-10, fixedcoupling=False, counts='weighted')),
 ("C3  crit x sub fc MLL", Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=True)),
 ("C4  pre x crit x sub fc", Integrand.pre_crit_sub('log', .1, 1e-10)),
]
for name, ig in cases:
    out = []
    for prec, n in (('f32', 10**9), ('f64', 10**8)):
        H.run(ig, 10**6, 1, e, prec); H.run(ig, n, 3, e, prec); torch.cuda.synchronize(); t = time.perf_counter()
        H.run(ig, n, 2, e, prec); torch.cuda.synchronize(); dt = time.perf_counter() - t
        out.append(f"{prec} {n/dt:.3e}/s")
    print(f"{name:40s} " + "   ".join(out))
