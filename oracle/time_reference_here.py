"""How fast is the oracle port next to the UNMODIFIED reference, on the same cores?

The reference is Python and cannot travel to the GPU box, so ``bench.py --impl reference`` times the oracle's
restatement there.  This script (build container only: needs /root/reference) runs the reference itself and the port
on the same C3 job -- criticalSampler + ungroomedSampler Python loops, C_groomed, twoEmissionWeight (running
coupling), np.nan_to_num, integrator.setBins/setDensity/integrate -- single-threaded, and records both rates in
profiles/reference_vs_port_cpu.json.  Usage: python oracle/time_reference_here.py [samples]"""
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300000
Z_CUT, BETA, EPS = .1, 2., 1e-10
edges = np.logspace(-11, np.log10(.5), 100)
np.seterr(all='ignore')

# ---- the port (what bench.py times on the GPU box); best of two runs each (the first pays for cold caches) ----
from oracle import jmc_oracle as O  # noqa: E402


def port_job():
    crit, jc = O.sample_loop_critical('log', n, Z_CUT, EPS, 1)
    sub, js = O.sample_loop_ungroomed('log', n, EPS, 2)
    csub = sub[:, 0] * sub[:, 1] ** BETA * sub[:, 1] ** BETA
    obs = np.maximum(O.ecf_groomed(crit[:, 0], crit[:, 1], Z_CUT, BETA, z_pre=Z_CUT, f=0., acc='MLL'), csub)
    w = np.nan_to_num(O.weight_two_emission(crit, sub, Z_CUT, BETA, 'quark', False)) * jc * js
    h, h2, c = O.histograms(obs, w, edges)
    dens, err = O.density_from_hist(h, h2, c, edges, np.log(1. / EPS) ** 4., n)
    O.cumulative(dens, err, edges, last_bc=[1., 'minus'])


def best_of_two(job):
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        job()
        times.append(time.perf_counter() - t0)
    return min(times)


t_port = best_of_two(port_job)

# ---- the reference ----
ref_shim.install()
from jetmontecarlo.montecarlo.integrator import integrator  # noqa: E402
from jetmontecarlo.numerics.radiators.samplers import criticalSampler, ungroomedSampler  # noqa: E402
from jetmontecarlo.numerics.observables import C_groomed  # noqa: E402
from jetmontecarlo.numerics.weights import twoEmissionWeight  # noqa: E402
random.seed(1)


def reference_job():
    cs = criticalSampler('log', zc=Z_CUT, epsilon=EPS)
    cs.generateSamples(n)
    us = ungroomedSampler('log', epsilon=EPS)
    us.generateSamples(n)
    crit_s, sub_s = cs.getSamples(), us.getSamples()
    c_sub = sub_s[:, 0] * sub_s[:, 1] ** BETA * sub_s[:, 1] ** BETA
    obs_r = np.maximum(C_groomed(crit_s[:, 0], crit_s[:, 1], Z_CUT, BETA, z_pre=Z_CUT, f=0., acc='MLL'), c_sub)
    w_r = np.nan_to_num(twoEmissionWeight(crit_s, sub_s, Z_CUT, BETA, 'quark', False))
    it = integrator()
    it.setLastBinBndCondition([1., 'minus'])
    it.bins, it.hasBins = edges, True
    it.setDensity(obs_r, w_r * np.array(cs.jacobians) * np.array(us.jacobians), cs.area * us.area)
    it.integrate()


t_ref = best_of_two(reference_job)

out = {"samples": n, "reference_samples_per_s": n / t_ref, "port_samples_per_s": n / t_port,
       "port_over_reference": t_ref / t_port, "threads": 1,
       "note": "same container, same core, same C3 job; the reference is the unmodified tree imported through "
               "oracle/ref_shim.py"}
print(json.dumps(out))
with open(os.path.join(os.path.dirname(HERE), "profiles", "reference_vs_port_cpu.json"), "w") as f:
    json.dump(out, f, indent=1)
