"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol
include/jmc_b200.h declares, the ctypes structs match the C layout, the host
mirror reproduces the reference's argument checks and areas, the product never
falls back to a CPU path, and the multi-GPU host logic works on gloo with
world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import jmc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "jmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from jetmontecarlo_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), name + " is declared in include/jmc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.jmc_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layout_matches_header():
    """compile a tiny C program against the header and compare sizeof/offsetof"""
    from jetmontecarlo_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "jmc_b200.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(jmc_sampler), sizeof(jmc_fn_params), sizeof(jmc_integrand));
  printf("%zu %zu %zu\n", offsetof(jmc_sampler, hi), offsetof(jmc_fn_params, s3), offsetof(jmc_integrand, shower_cutoff));
  printf("%zu %zu\n", offsetof(jmc_integrand, epsilon), offsetof(jmc_fn_params, z_cut));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o",
                               os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [int(x) for x in out]
    want = [C.sizeof(_lib.Sampler), C.sizeof(_lib.FnParams), C.sizeof(_lib.Integrand),
            _lib.Sampler.hi.offset, _lib.FnParams.s3.offset, _lib.Integrand.shower_cutoff.offset,
            _lib.Integrand.epsilon.offset, _lib.FnParams.z_cut.offset]
    assert got == want


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point raises instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the no-GPU behaviour is checked on the CPU box")
    from jetmontecarlo_b200.montecarlo.integrator import integrator
    from jetmontecarlo_b200.numerics.observables import C_ungroomed
    from jetmontecarlo_b200.numerics.radiators.samplers import ungroomedSampler
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        C_ungroomed(np.ones(3), np.ones(3), 2.)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ungroomedSampler('lin').generateSamples(4)
    it = integrator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        it.setBins(10, np.ones(4), 'lin')
    # and the product package never imports the oracle
    pkg = os.path.join(ROOT, "jetmontecarlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|#include\s+[\"<][^\n]*oracle|import_module\([^)]*oracle",
                                     text, flags=re.M), os.path.join(dirpath, f)


def test_missing_library_fails_loudly(tmp_path):
    """import of the package without libjmc_b200.so raises ImportError"""
    code = ("import sys, importlib.util, os\n"
            "sys.path.insert(0, %r)\n"
            "import jetmontecarlo_b200._lib as L\n" % ROOT)
    env = dict(os.environ)
    # run in a copy of the package that has no lib/ directory
    import shutil
    dst = tmp_path / "jetmontecarlo_b200"
    shutil.copytree(os.path.join(ROOT, "jetmontecarlo_b200"), dst,
                    ignore=shutil.ignore_patterns("lib", "build", "__pycache__"))
    res = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import jetmontecarlo_b200"
                          % str(tmp_path)], capture_output=True, text=True, env=env)
    assert res.returncode != 0 and "no CPU fallback" in res.stderr


def test_sampler_argument_checks_and_areas():
    """same assertions and messages as the reference (sampler.py:65-81, samplers.py:12-18,116,298;
    areas: simple_tests/test_jetSamplers.py:50,83,120,151)"""
    from jetmontecarlo_b200.montecarlo.sampler import simpleSampler
    from jetmontecarlo_b200.numerics.radiators.samplers import (criticalSampler, precriticalSampler,
                                                                ungroomedSampler)
    with pytest.raises(AssertionError, match="is not a supported sample method"):
        ungroomedSampler('cubic')
    for eps in (0., 1., -1., None, 2.):
        with pytest.raises(AssertionError, match="is invalid for log sampling"):
            criticalSampler('log', zc=.1, epsilon=eps)
    for zc in (0., .5, .7, -1, None):
        with pytest.raises(AssertionError, match="invalid grooming parameter"):
            criticalSampler('lin', zc=zc)
        with pytest.raises(AssertionError, match="invalid grooming parameter"):
            precriticalSampler('lin', zc=zc)
    with pytest.raises(AssertionError, match="radius must be greater than zero"):
        ungroomedSampler('lin', radius=0.)
    assert criticalSampler('lin', zc=.1, radius=.8).area == (1. / 2. - .1) * .8
    assert criticalSampler('log', zc=.1, epsilon=1e-10).area == np.log(1. / 1e-10) ** 2.
    assert precriticalSampler('lin', zc=.2).area == .2
    assert precriticalSampler('log', zc=.2, epsilon=1e-5).area == np.log(1. / 1e-5)
    assert ungroomedSampler('lin', radius=.6).area == .6 / 2.
    assert ungroomedSampler('log', epsilon=1e-3).area == np.log(1. / 1e-3) ** 2.
    assert simpleSampler('lin', bounds=[.1, .5]).area == .5 - .1
    assert simpleSampler('log', epsilon=1e-8).area == np.log(1. / 1e-8)
    s = criticalSampler('lin', zc=.1)
    assert s.samples == [] and s.jacobians == [] and s.getSampleMethod() == 'lin'
    # C-side area agrees with the NumPy one
    from jetmontecarlo_b200 import _lib
    a = C.c_double()
    d = _lib.Sampler(kind=_lib.SAMPLER_CRITICAL, method=_lib.LOG, epsilon=1e-10, zc=.1, radius=1.)
    assert _lib.lib.jmc_sampler_area(C.byref(d), C.byref(a)) == 0 and a.value == s.__class__('log', .1, 1e-10).area


def test_integrand_descriptions():
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.fused import Integrand
    g = Integrand.crit_sub('log', .1, 1e-10, fixedcoupling=False)
    assert g.n_uniform == 4 and g.z_pre == .1 and g.f == 0. and g.nan_to_num
    l2 = np.log(1. / 1e-10) ** 2.
    assert g.area == l2 * l2
    a = C.c_double()
    s = g.c_struct()
    assert _lib.lib.jmc_integrand_area(C.byref(s), C.byref(a)) == 0 and abs(a.value / g.area - 1) < 1e-15
    assert Integrand.pre_crit_sub('log', .1, 1e-10).n_uniform == 6
    assert Integrand.radiator('lin', radius=.8, theta_scale=.5).area == .8 / 2. * .5
    with pytest.raises(TypeError):
        Integrand.pre_crit_sub('log', .1, 1e-10, fixedcoupling=False)
    with pytest.raises(AssertionError):
        Integrand.crit_sub('log', .6, 1e-10)
    with pytest.raises(AssertionError):
        Integrand.radiator('log', epsilon=0.)
    # C-side validation returns JMC_EINVAL with the reference's wording
    bad = g.c_struct()
    bad.z_cut = .7
    rc = _lib.lib.jmc_run(C.byref(bad), 10, 0, 0, 0, None, 0, None, None, None, None, None)
    assert rc == -1 and b"invalid grooming parameter" in _lib.lib.jmc_last_error()
    with pytest.raises(_lib.JmcError):
        _lib.check(rc)


def test_host_helpers_match_oracle(golden):
    from jetmontecarlo_b200.analytics import qcd_utils as Q
    from jetmontecarlo_b200.numerics.observables import C_ungroomed_max
    from jetmontecarlo_b200.utils.interpolation import lin_log_mixed_list
    assert (Q.CF, Q.CA, Q.TF, Q.N_F, Q.beta_0, Q.alpha_fixed, Q.MU_NP) == (
        O.CF, O.CA, O.TF, O.N_F, O.BETA_0, O.ALPHA_FIXED, O.MU_NP)
    assert Q.CR('quark') == O.CF and Q.CR('gluon') == O.CA
    with pytest.raises(AssertionError, match="is not a supported jet type"):
        Q.CR('photon')
    g = golden("lin_log_mixed")
    np.testing.assert_array_equal(lin_log_mixed_list(1e-10, 1., 12), g['a'])
    np.testing.assert_array_equal(lin_log_mixed_list(1e-15, 1., 501), g['b'])
    assert C_ungroomed_max(2., .5, 'LL') == O.ecf_ungroomed_max(2., .5, 'LL')
    assert C_ungroomed_max(3., .5, 'MLL') == O.ecf_ungroomed_max(3., .5, 'MLL')


def test_shard_bounds_cover_the_job():
    from jetmontecarlo_b200.distributed import shard_bounds
    for n in (0, 1, 7, 10 ** 10, 10 ** 9 + 3):
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                first, cnt = shard_bounds(n, r, world)
                assert first == nxt and cnt in (n // world, n // world + 1)
                nxt += cnt
            assert nxt == n


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from jetmontecarlo_b200 import distributed as du
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2)
first, cnt = du.shard(1001)
assert (first, cnt) == ((0, 501) if rank == 0 else (501, 500))
sw = torch.arange(5, dtype=torch.float64) * (rank + 1)
sw2 = torch.ones(5, dtype=torch.float64) * (rank + 1)
cnt_t = torch.tensor([1, 2, 3, 2 ** 40, 0], dtype=torch.int64) * (rank + 1)
du.allreduce_histograms(sw, sw2, cnt_t)
assert torch.equal(sw, torch.arange(5, dtype=torch.float64) * 3)
assert torch.equal(sw2, torch.full((5,), 3., dtype=torch.float64))
assert torch.equal(cnt_t, torch.tensor([3, 6, 9, 3 * 2 ** 40, 0], dtype=torch.int64))
mm = torch.tensor([0.5 + rank, 7. - rank], dtype=torch.float64)
du.allreduce_minmax(mm)
assert mm.tolist() == [0.5, 7.0]
mm = torch.tensor([float("nan") if rank == 1 else 1., 2.], dtype=torch.float64)
du.allreduce_minmax(mm)
assert torch.isnan(mm).all()
# a rank without samples contributes [inf, -inf]
mm = torch.tensor([float("inf"), float("-inf")] if rank == 0 else [3., 4.], dtype=torch.float64)
du.allreduce_minmax(mm)
assert mm.tolist() == [3.0, 4.0]
mm = torch.tensor([float("inf"), float("-inf")], dtype=torch.float64)
du.allreduce_minmax(mm)
assert mm.tolist() == [float("inf"), float("-inf")]
# a whole parameter grid in one collective: rows are independent, NaN stays in its row
grid = torch.tensor([[1. + rank, 5. + rank], [2. - rank, 9. - rank], [float("nan") if rank == 0 else 0., 1.],
                     [-3., -2. * (rank + 1)]], dtype=torch.float64)
du.allreduce_minmax(grid)
assert grid[[0, 1, 3]].tolist() == [[1., 6.], [1., 9.], [-3., -2.]]
assert torch.isnan(grid[2]).all()
# a parameter grid sharded by rows: every rank contributes its block of finished rows, one all-gather, no reduction
for n_total in (5, 4, 1):
    p0, cnt = du.shard_bounds(n_total, rank, 2)
    rows = (torch.arange(n_total, dtype=torch.float64)[:, None] * 10 + torch.arange(3, dtype=torch.float64)[None, :])[p0:p0 + cnt]
    full = du.gather_rows(rows, n_total, 2)
    assert full.shape == (n_total, 3) and torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float64) * 10)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gloo_world_size_2():
    """sharding + the single histogram all-reduce + min/max merge on two CPU ranks"""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    code = _GLOO_WORKER % dict(root=ROOT, port=port)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err
        assert "ok" in out


def test_c_abi_argument_errors_need_no_device():
    """argument validation comes first: these calls fail with JMC_EINVAL and a message before any CUDA call, so they
    behave the same with and without a GPU"""
    from jetmontecarlo_b200 import _lib
    from jetmontecarlo_b200.fused import Integrand
    L = _lib.lib
    s = Integrand.crit_sub('log', .1, 1e-10).c_struct()
    buf = np.zeros(8)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    cases = [
        (L.jmc_run(None, 10, 0, 0, 0, None, 0, None, None, None, None, None), b"integrand is NULL"),
        (L.jmc_run(C.byref(s), 10, 0, 0, 7, None, 0, None, None, None, None, None), b"precision must be"),
        (L.jmc_run_batch(C.byref(s), 0, 10, 0, 0, 1, None, 0, None, None, None, None, None), b"at least one parameter point"),
        (L.jmc_dump(None, 10, 0, 0, 0, None, None, None, None), b"integrand is NULL"),
        (L.jmc_set_density_host(P(buf), P(buf), 8, P(buf), 1, 1., 0, 0., None, None, None, None, None, None, None),
         b"Need bins to produce density histogram."),
        (L.jmc_set_density_host(None, None, 8, P(buf), 8, 1., 0, 0., None, None, None, None, None, None, None),
         b"NULL input array"),
        (L.jmc_integrate_host(C.byref(s), 10, 0, 0, 1, None, 0, 0, 0., None, None, None, None, None, None, None),
         b"Need bins to produce density histogram."),
        (L.jmc_sample_host(None, 4, 0, 0, P(buf), P(buf)), b"sampler is NULL"),
        (L.jmc_sampler_area(None, None), b"NULL argument"),
    ]
    for rc, text in cases:
        assert rc == -1, (rc, text)
    # the message of the LAST failing call is the one kept (per thread)
    assert b"NULL argument" in L.jmc_last_error()
    bad = Integrand.crit_sub('log', .1, 1e-10).c_struct()
    for field, value, text in (("kind", 99, b"unknown integrand kind"), ("method", 5, b"not a supported sample method"),
                               ("jet", 3, b"not a supported jet type"), ("epsilon", 0., b"invalid for log sampling"),
                               ("radius", 0., b"radius must be greater than zero"), ("obs_acc", 9, b"Accuracy must be LL or MLL")):
        g = Integrand.crit_sub('log', .1, 1e-10).c_struct()
        setattr(g, field, value)
        assert L.jmc_run(C.byref(g), 10, 0, 0, 0, None, 0, None, None, None, None, None) == -1
        assert text in L.jmc_last_error(), (field, L.jmc_last_error())


def test_header_is_c_and_library_is_callable_from_c(tmp_path):
    """include/jmc_b200.h compiles as plain C99 and a C program can bind the library through it (what a non-Python
    host of the reference's path would do)"""
    import shutil
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not installed")
    header = os.path.join(ROOT, "include", "jmc_b200.h")
    res = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include <stdio.h>
#include <dlfcn.h>
#include "jmc_b200.h"
int main(int argc, char** argv) {
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }
    int (*ver)(void) = (int (*)(void))dlsym(h, "jmc_abi_version");
    int (*area)(const jmc_sampler*, double*) = (int (*)(const jmc_sampler*, double*))dlsym(h, "jmc_sampler_area");
    const char* (*err)(void) = (const char* (*)(void))dlsym(h, "jmc_last_error");
    jmc_sampler s = {0};
    double a = 0;
    s.kind = JMC_SAMPLER_CRITICAL; s.method = JMC_LIN; s.zc = .1; s.radius = .8;
    int rc = area(&s, &a);
    printf("abi %d rc %d area %.17g\n", ver(), rc, a);
    rc = area(0, 0);
    printf("rc %d err %s\n", rc, err());
    return 0;
}
''')
    exe = tmp_path / "probe"
    res = subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-ldl"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = os.path.join(ROOT, "jetmontecarlo_b200", "lib", "libjmc_b200.so")
    out = subprocess.run([str(exe), lib], capture_output=True, text=True).stdout.splitlines()
    assert out[0] == "abi 2 rc 0 area %.17g" % ((1. / 2. - .1) * .8)
    assert out[1] == "rc -1 err jmc_sampler_area: NULL argument"
