"""Multi-GPU: the sharded job equals the single-GPU job (same sample indices; counts exact).  Needs >= 2 GPUs;
on a single-GPU box the same code path is covered by two ranks sharing cuda:0 over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ngpu = torch.cuda.device_count()
torch.cuda.set_device(rank %% ngpu)
backend = "nccl" if ngpu >= world else "gloo"
dist.init_process_group(backend, init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
from jetmontecarlo_b200 import fused, distributed as du
from jetmontecarlo_b200.fused import Integrand
if backend == "gloo":            # gloo reduces host tensors: route the packed histogram through the host
    _ar = dist.all_reduce
    def all_reduce(t, op=dist.ReduceOp.SUM, group=None):
        h = t.cpu(); _ar(h, op=op, group=group); t.copy_(h)
    dist.all_reduce = all_reduce
ig = Integrand.crit_sub('log', .1, 1e-10, jet_type='quark', fixedcoupling=False, obs_acc='MLL')
edges = np.logspace(-11, np.log10(.5), 100)
n = 20_000_001
it = fused.integrate(ig, n, edges=edges, seed=3, precision='f32', last_bc=[1., 'minus'])
it2 = fused.integrate(ig, n, numbins=50, binspacing='log', seed=3, precision='f64', last_bc=[1., 'minus'])
# BASELINE config 5: the parameter GRID is sharded (rows gathered, nothing reduced)
if backend == "gloo":
    _ag = dist.all_gather_into_tensor
    def all_gather_into_tensor(out, inp, group=None):
        ho, hi = out.cpu(), inp.cpu(); _ag(ho, hi, group=group); out.copy_(ho)
    dist.all_gather_into_tensor = all_gather_into_tensor
grid = np.array([1e-3, 1e-2, .1, .3, 1.])
igs = [Integrand.radiator('log', 1e-8, beta=2., jet_type='quark', fixedcoupling=True, theta_scale=float(t)) for t in grid]
res = fused.integrate_batch(igs, 400_000, numbins=30, binspacing='log', min_log_bins=grid ** 2. * 1e-15, seed=4,
                            precision='f32', last_bc=[0., 'plus'])
if rank == 0:
    np.savez(%(out)r, count=it._count.cpu().numpy(), sum_w=it._sum_w.cpu().numpy(), integral=np.asarray(it.integral),
             bins2=it2.bins, count2=it2._count.cpu().numpy(), backend=backend, grid_bins=res['bins'],
             grid_count=res['count'], grid_integral=res['integral'])
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_equal_one(cuda, tmp_path):
    import torch
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.fused import Integrand
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "multi.npz")
    code = _WORKER % dict(root=ROOT, port=port, out=out)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        _, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-3000:]
    got = np.load(out)
    ig = Integrand.crit_sub('log', .1, 1e-10, jet_type='quark', fixedcoupling=False, obs_acc='MLL')
    edges = np.logspace(-11, np.log10(.5), 100)
    n = 20_000_001
    one = fused.integrate(ig, n, edges=edges, seed=3, precision='f32', last_bc=[1., 'minus'])
    np.testing.assert_array_equal(got['count'], one._count.cpu().numpy())
    np.testing.assert_allclose(got['sum_w'], one._sum_w.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(got['integral'], np.asarray(one.integral), rtol=2e-6, atol=1e-9)
    one2 = fused.integrate(ig, n, numbins=50, binspacing='log', seed=3, precision='f64', last_bc=[1., 'minus'])
    np.testing.assert_array_equal(got['bins2'], one2.bins)          # data-dependent edges via the min/max all-reduce
    np.testing.assert_array_equal(got['count2'], one2._count.cpu().numpy())
    # the sharded grid (rank 0: points 0-2, rank 1: points 3-4) equals the grid on one GPU
    grid = np.array([1e-3, 1e-2, .1, .3, 1.])
    igs = [Integrand.radiator('log', 1e-8, beta=2., jet_type='quark', fixedcoupling=True, theta_scale=float(t)) for t in grid]
    res = fused.integrate_batch(igs, 400_000, numbins=30, binspacing='log', min_log_bins=grid ** 2. * 1e-15, seed=4,
                                precision='f32', last_bc=[0., 'plus'])
    np.testing.assert_array_equal(got['grid_bins'], res['bins'])
    np.testing.assert_array_equal(got['grid_count'], res['count'])
    np.testing.assert_allclose(got['grid_integral'], res['integral'], rtol=1e-9)
