"""Replay (array) path at 1e7 samples: the reference's script flow through the mirror API, host arrays vs
device-resident tensors, per-stage times.  Also runs the shower so that one ncu pass covers all kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jetmontecarlo_b200.montecarlo.sampler import set_seed
from jetmontecarlo_b200.montecarlo.integrator import integrator
from jetmontecarlo_b200.numerics.radiators.samplers import criticalSampler, ungroomedSampler
from jetmontecarlo_b200.numerics.observables import C_groomed
from jetmontecarlo_b200.numerics.weights import twoEmissionWeight, criticalEmissionWeight
from jetmontecarlo_b200._arrayfn import eval_fn
from jetmontecarlo_b200 import _lib, fused
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
set_seed(3)
def T(label, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
    dt = time.perf_counter() - t; print(f"  {label:58s} {dt*1e3:9.2f} ms  {n/dt:.3e} samples/s"); return r
for rep in range(2):
    print("pass", rep, "(pass 0 includes CUDA start-up)")
    cs = criticalSampler('log', zc=.1, epsilon=1e-10); us = ungroomedSampler('log', epsilon=1e-10)
    T("criticalSampler.generateSamples (samples -> host)", lambda: cs.generateSamples(n))
    T("ungroomedSampler.generateSamples", lambda: us.generateSamples(n))
    c, s = cs.device_samples, us.device_samples
    obs = T("C_groomed + csub + max (device tensors)", lambda: torch.maximum(
        C_groomed(c[:, 0].contiguous(), c[:, 1].contiguous(), .1, 2., z_pre=.1, f=0., acc='MLL'),
        eval_fn(_lib.FN_MUL, [eval_fn(_lib.FN_MUL, [s[:, 0].contiguous(), s[:, 1].contiguous() ** 2]), s[:, 1].contiguous() ** 2])))
    w = T("twoEmissionWeight rc (device tensors)", lambda: twoEmissionWeight(c, s, .1, 2., 'quark', False))
    wj = T("nan_to_num * jacobians", lambda: torch.nan_to_num(w) * cs.device_jacobians * us.device_jacobians)
    it = integrator(); it.setLastBinBndCondition([1., 'minus'])
    T("integrator.setBins (min/max)", lambda: it.setBins(100, obs, 'log'))
    T("integrator.setDensity (hist + density)", lambda: it.setDensity(obs, wj, cs.area * us.area))
    T("integrator.integrate", lambda: it.integrate())
    oh, wh = obs.cpu().numpy(), wj.cpu().numpy()
    it2 = integrator(); it2.setLastBinBndCondition([1., 'minus']); it2.bins, it2.hasBins = it.bins, True
    T("setDensity from HOST arrays (incl. 16 B/sample over PCIe)", lambda: it2.setDensity(oh, wh, cs.area * us.area))
ig = fused.Integrand.shower(beta=2., jet_type='quark', acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL', n_emissions=2)
edges = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
for prec in ('f32', 'f64'):
    T("shower MLL quark softdrop " + prec + " (jets)", lambda: fused.integrate(ig, n, edges=edges, seed=1, precision=prec, first_bc=0.))
