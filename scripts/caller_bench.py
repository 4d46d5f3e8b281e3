"""Time the batched callers of the hot path at the reference's default sizes (examples/params.py:67,86:
NUM_MC_EVENTS = 5e6, NUM_RAD_BINS = 5e3), on a reduced grid, and extrapolate."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jetmontecarlo_b200.montecarlo.sampler import set_seed
from jetmontecarlo_b200.numerics.radiators.samplers import ungroomedSampler, criticalSampler
from jetmontecarlo_b200.numerics.radiators import generation as GEN
from jetmontecarlo_b200.numerics.splitting import gen_normalized_splitting
from jetmontecarlo_b200 import _lib
set_seed(1)
n = 5_000_000
t = time.perf_counter(); s = ungroomedSampler('log', epsilon=1e-10); s.generateSamples(n); torch.cuda.synchronize()
print(f"ungroomedSampler.generateSamples({n:.0e}) incl. copy of samples to host: {time.perf_counter()-t:.3f} s  (reference: 49 s)")
for grid in (40, 80):
    l0 = _lib.launch_count(); torch.cuda.synchronize(); t = time.perf_counter()
    fn, d = GEN.gen_crit_sub_num_rad(s, jet_type='quark', obs_accuracy='MLL', splitfn_accuracy='MLL', beta=2., epsilon=1e-10,
                                     bin_space='log', fixed_coupling=False, num_bins=grid)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    pts = len(set(d['ys']))
    print(f"gen_crit_sub_num_rad: {pts} grid points x {n:.0e} samples, {grid} bins: {dt:.3f} s = {dt/pts*1e3:.2f} ms/point, "
          f"{pts*n/dt:.3e} sample-evals/s, {(_lib.launch_count()-l0)/pts:.1f} launches/point -> 5e3 points: {dt/pts*5e3:.0f} s")
t = time.perf_counter()
fn = gen_normalized_splitting(n, .1, jet_type='quark', accuracy='MLL', fixed_coupling=False, bin_space='lin', epsilon=1e-10, num_bins=40)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"gen_normalized_splitting: 40 grid points x {n:.0e} fresh samples: {dt:.3f} s = {dt/40*1e3:.2f} ms/point -> 5e3 points: {dt/40*5e3:.0f} s")
