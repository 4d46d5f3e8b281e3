"""Fused path: describe an integrand, and the GPU generates, weights and bins
its samples in one kernel (jmc_run) without ever writing a sample to HBM.

An ``Integrand`` names one of the sampler x observable x weight combinations
the reference's scripts build by hand:

  Integrand.radiator(...)      ungroomedSampler, C_ungroomed, radiatorWeight          (C1, C2, C5)
  Integrand.critical(...)      criticalSampler, C_groomed, criticalEmissionWeight
  Integrand.crit_sub(...)      critical x ungroomed, max(C_groomed, csub), twoEmissionWeight   (C3)
  Integrand.pre_crit_sub(...)  + precriticalSampler with the zero-bin draw            (C4)
  Integrand.shower(...)        veto-algorithm angularity shower, ECF of the chain     (C2, C3)

``integrate(integrand, n, edges | (numbins, spacing), ...)`` returns a filled
``integrator``.  With ``torch.distributed`` initialised the sample indices are
sharded over the ranks and the histograms are summed by one all-reduce.
"""
import ctypes as C

import numpy as np
import torch

from . import _device as D
from . import _lib
from . import distributed as dist_utils

PRECISIONS = {'f64': _lib.F64, 'f32': _lib.F32}


class Integrand:
    """Plain description of an integrand (mirrors ``jmc_integrand``)."""

    def __init__(self, kind, method, *, jet_type='quark', fixedcoupling=True, obs_acc='LL',
                 split_acc='LL', nan_to_num=False, epsilon=0., z_cut=0., radius=1., beta=2., f=1.,
                 z_pre=0., theta_scale=1., bounds=(0., 1.), groomer=0, n_emissions=1, beta_sd=0.,
                 cutoff=0., counts='all'):
        assert method in ('lin', 'log'), str(method) + "is not a supported sample method."
        assert jet_type in ('quark', 'gluon'), str(jet_type) + "is not a supported jet type."
        assert obs_acc in ('LL', 'MLL'), "Accuracy must be LL or MLL"
        assert split_acc in ('LL', 'LL_nonsing', 'MLL'), "Invalid accuracy."
        eps = 0 if epsilon is None else epsilon
        if method == 'log':
            assert 0 < eps < 1., ("epsilon={epsilon} is invalid for {type} sampling."
                                  .format(epsilon=eps, type=method))
        self.kind, self.method = kind, method
        self.jet_type, self.fixedcoupling = jet_type, bool(fixedcoupling)
        self.obs_acc, self.split_acc, self.nan_to_num = obs_acc, split_acc, bool(nan_to_num)
        self.epsilon, self.z_cut, self.radius, self.beta = float(eps), float(z_cut), float(radius), float(beta)
        self.f, self.z_pre, self.theta_scale = float(f), float(z_pre), float(theta_scale)
        self.bounds = (float(bounds[0]), float(bounds[1]))
        self.groomer, self.beta_sd, self.cutoff = int(groomer), float(beta_sd), float(cutoff)
        self.n_emissions = 0 if n_emissions == 'all' else int(n_emissions)
        self.rss_flags = 0
        # 'all': per-bin counts as np.histogram gives them; 'weighted': only samples with a non-zero weight are
        # counted (same density / integral; lets the FP32 path skip the samples whose weight is NaN -> 0)
        assert counts in ('all', 'weighted'), "counts must be 'all' or 'weighted'"
        self.counts = counts

    # ---- constructors ---------------------------------------------------------
    @classmethod
    def radiator(cls, method, epsilon=0., radius=1., beta=2., jet_type='quark', fixedcoupling=True,
                 obs_acc='LL', split_acc='LL', theta_scale=1., nan_to_num=False):
        assert radius > 0, "Jet radius must be greater than zero!"
        return cls(_lib.KIND_RADIATOR, method, epsilon=epsilon, radius=radius, beta=beta, jet_type=jet_type,
                   fixedcoupling=fixedcoupling, obs_acc=obs_acc, split_acc=split_acc, theta_scale=theta_scale,
                   nan_to_num=nan_to_num)

    @classmethod
    def critical(cls, method, z_cut, epsilon=0., radius=1., beta=2., jet_type='quark', fixedcoupling=True,
                 obs_acc='LL', z_pre=0., f=1., nan_to_num=None, counts='all'):
        _check_zcut(z_cut)
        nan_to_num = (not fixedcoupling) if nan_to_num is None else nan_to_num
        return cls(_lib.KIND_CRIT, method, epsilon=epsilon, z_cut=z_cut, radius=radius, beta=beta,
                   jet_type=jet_type, fixedcoupling=fixedcoupling, obs_acc=obs_acc, z_pre=z_pre, f=f,
                   nan_to_num=nan_to_num, counts=counts)

    @classmethod
    def crit_sub(cls, method, z_cut, epsilon=0., radius=1., beta=2., jet_type='quark', fixedcoupling=True,
                 obs_acc='MLL', z_pre=None, f=0., nan_to_num=True, counts='all'):
        """Defaults are the Soft Drop / mMDT form ``z_pre=z_cut, f=0``
        (examples/sudakov_comparisons/sudakov_utils.py:481-486)."""
        _check_zcut(z_cut)
        return cls(_lib.KIND_CRIT_SUB, method, epsilon=epsilon, z_cut=z_cut, radius=radius, beta=beta,
                   jet_type=jet_type, fixedcoupling=fixedcoupling, obs_acc=obs_acc,
                   z_pre=z_cut if z_pre is None else z_pre, f=f, nan_to_num=nan_to_num, counts=counts)

    @classmethod
    def pre_crit_sub(cls, method, z_cut, epsilon=0., radius=1., beta=2., jet_type='quark', fixedcoupling=True,
                     obs_acc='LL', f=1., nan_to_num=False):
        _check_zcut(z_cut)
        if not fixedcoupling:
            raise TypeError("Can only accept fixed coupling for now!")
        assert method == 'log', "Weight requires valid cutoff for zero bin."
        return cls(_lib.KIND_PRE_CRIT_SUB, method, epsilon=epsilon, z_cut=z_cut, radius=radius, beta=beta,
                   jet_type=jet_type, fixedcoupling=True, obs_acc=obs_acc, f=f, nan_to_num=nan_to_num)

    @classmethod
    def simple(cls, method, bounds=(0., 1.), epsilon=0.):
        return cls(_lib.KIND_SIMPLE, method, epsilon=epsilon, bounds=bounds)

    @classmethod
    def shower(cls, beta=2., jet_type='quark', acc='MLL', radius=1., cutoff=None, groomer='ungroomed',
               z_cut=0., beta_sd=0., obs_acc='LL', n_emissions=1, f=1., emission_type='precritsub', few_pres=True):
        """``groomer``: 'ungroomed', 'softdrop' (z_cut, beta_sd) or 'rss' (recursive safe subtraction: z_cut, f,
        emission_type matched by substring as the reference does, few_pres; utils/partonshower_utils.py:501-674)."""
        from .analytics.qcd_utils import MU_NP
        assert acc in ('LL', 'MLL'), "Invalid accuracy."
        assert groomer in ('ungroomed', 'softdrop', 'rss')
        if groomer == 'rss':
            assert f > 1 / 3, "f = " + str(f) + " is less than 1/3, and the grooming may groom away harder branches first."
        ig = cls(_lib.KIND_SHOWER, 'lin', radius=radius, beta=beta, jet_type=jet_type, split_acc=acc,
                 obs_acc=obs_acc, z_cut=z_cut, groomer={'ungroomed': 0, 'softdrop': 1, 'rss': 2}[groomer],
                 n_emissions=n_emissions, beta_sd=beta_sd, cutoff=MU_NP if cutoff is None else cutoff, f=f)
        ig.rss_flags = ((_lib.RSS_PRECRIT if 'precrit' in emission_type else 0)
                        | (_lib.RSS_CRIT if 'crit' in emission_type else 0)
                        | (_lib.RSS_SUB if 'sub' in emission_type else 0) | (_lib.RSS_FEW_PRES if few_pres else 0))
        return ig

    # ---- C struct and derived quantities ---------------------------------------
    def _c_values(self):
        """the fields of jmc_integrand in declaration order (include/jmc_b200.h, _lib.Integrand)"""
        return (self.kind, _lib.METHODS[self.method], _lib.JETS[self.jet_type], int(self.fixedcoupling),
                _lib.ACCS[self.obs_acc], _lib.ACCS[self.split_acc], int(self.nan_to_num), self.groomer,
                self.n_emissions, int(self.counts == 'weighted'), self.epsilon, self.z_cut, self.radius, self.beta,
                self.f, self.z_pre, self.theta_scale, self.bounds[0], self.bounds[1], self.beta_sd, self.cutoff,
                self.rss_flags, 0)

    def c_struct(self):
        return _lib.Integrand(*self._c_values())

    @property
    def area(self):
        """Product of the sampler areas, as the reference's scripts multiply them
        (NumPy arithmetic, so it is bit-identical to ``sampler.area`` products)."""
        if self.kind == _lib.KIND_SHOWER:
            return 1.
        if self.method == 'log':
            l1 = np.log(1. / self.epsilon)
            l2 = np.log(1. / self.epsilon) ** 2.
            return {_lib.KIND_RADIATOR: l2, _lib.KIND_CRIT: l2, _lib.KIND_CRIT_SUB: l2 * l2,
                    _lib.KIND_PRE_CRIT_SUB: l2 * l1 * l2, _lib.KIND_SIMPLE: l1}[self.kind]
        crit = (1. / 2. - self.z_cut) * self.radius
        ung = self.radius / 2.
        return {_lib.KIND_RADIATOR: ung * self.theta_scale, _lib.KIND_CRIT: crit,
                _lib.KIND_CRIT_SUB: crit * ung, _lib.KIND_PRE_CRIT_SUB: crit * self.z_cut * ung,
                _lib.KIND_SIMPLE: self.bounds[1] - self.bounds[0]}[self.kind]

    @property
    def n_uniform(self):
        s = self.c_struct()
        return _lib.check(_lib.lib.jmc_integrand_uniforms(C.byref(s)))


def _check_zcut(zc):
    zc = 0 if zc is None else zc
    assert 0 < zc < 1. / 2., ("zc={zcut} is an invalid grooming parameter.".format(zcut=zc))


def accumulate(integrand, num_samples, seed, precision, edges_dev, sum_w, sum_w2, count, group=None,
               minmax=None, edges_host=None):
    """Run this rank's share of ``num_samples`` and (if distributed) all-reduce.

    The tensors are CUDA tensors: edges float64 [n_edges], sum_w / sum_w2
    float64 [n_edges-1], count int64 [n_edges-1]; they are accumulated into.
    With ``edges_host`` (the same edges as a NumPy array) the job goes through the calling thread's ``jmc_handle``:
    nothing is read back, nothing synchronises, and a repeated job re-derives nothing."""
    D.require_cuda()
    num_samples = int(num_samples)
    first, n_local = dist_utils.shard(num_samples, group)
    s = integrand.c_struct()
    n_edges = 0 if edges_dev is None else int(edges_dev.numel())
    key, prec = int(seed) & 0xFFFFFFFFFFFFFFFF, PRECISIONS[precision]
    if edges_host is not None or edges_dev is None:
        eh = None if edges_dev is None else np.ascontiguousarray(edges_host, dtype=np.float64)
        handle = _lib.default_handle(torch.cuda.current_device())
        _lib.check(_lib.lib.jmc_handle_run(handle.ptr, C.byref(s), n_local, key, first, prec,
                                           None if eh is None else eh.ctypes.data_as(C.c_void_p), n_edges,
                                           D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), D.ptr(minmax), D.stream()))
    else:
        _lib.check(_lib.lib.jmc_run(C.byref(s), n_local, key, first, prec, D.ptr(edges_dev), n_edges, D.ptr(sum_w),
                                    D.ptr(sum_w2), D.ptr(count), D.ptr(minmax), D.stream()))
    if edges_dev is not None:
        dist_utils.allreduce_histograms(sum_w, sum_w2, count, group)
    if minmax is not None:
        dist_utils.allreduce_minmax(minmax, group)


def observable_range(integrand, num_samples, seed, precision='f64', group=None):
    """(min, max) of the observable over the job's samples, regenerated from the
    counter RNG (nothing stored) -- the data-dependent half of setBins."""
    mm = torch.tensor([np.inf, -np.inf], dtype=torch.float64, device=D.device())
    accumulate(integrand, num_samples, seed, precision, None, None, None, None, group, minmax=mm)
    lo, hi = D.to_host(mm)
    return lo, hi


def integrate(integrand, num_samples, *, edges=None, numbins=None, binspacing='log', seed=0, precision='f32',
              last_bc=None, first_bc=None, min_log_bin=1e-15, group=None):
    """One call for the whole job; returns an ``integrator`` with density,
    densityErr, integral and integralErr filled."""
    from .montecarlo.integrator import integrator
    it = integrator()
    if last_bc is not None:
        it.setLastBinBndCondition(list(last_bc))
    if first_bc is not None:
        it.setFirstBinBndCondition(first_bc)
    if edges is not None:
        it.bins = np.asarray(edges, dtype=np.float64)
        it.binspacing = binspacing
        with np.errstate(all='ignore'):
            it.bin_midpoints = ((it.bins[1:] + it.bins[:-1]) / 2. if binspacing == 'lin'
                                else np.exp((np.log(it.bins[1:]) + np.log(it.bins[:-1])) / 2.))
        it.hasBins = True
    else:
        # same arithmetic mode for both passes: the min/max pass must see the samples the histogram pass bins
        it.setBinsFused(numbins, integrand, num_samples, binspacing, seed, min_log_bin=min_log_bin,
                        precision=precision, group=group)
    it.integrateFused(integrand, num_samples, seed, precision=precision, group=group)
    return it


def _as_struct_array(integrands):
    """contiguous jmc_integrand[len(integrands)] for the batched entry points: the rows are assembled by NumPy from
    the fields' tuples (a 5000-point grid costs 10 ms instead of 50 ms of per-element ctypes assignments); the
    returned object keeps the memory alive and converts to the pointer ctypes expects"""
    table = np.array([ig._c_values() for ig in integrands], dtype=np.dtype(_lib.Integrand))
    arr = (_lib.Integrand * len(integrands)).from_buffer(table)
    arr._table = table
    return arr


def _set_bins_rows(lo, hi, min_log_bins, numbins, binspacing):
    """integrator.setBins for every row of a grid (integrator.py:35-58): ``np.linspace(min(0, lo), hi, numbins)`` or
    ``np.logspace(log10(max(lo, min_log_bin)), log10(hi), numbins)`` per point -- NumPy's own array form of the same
    calls (identical bits), in row blocks on the host's cores (np.power releases the GIL; 5000 x 5000 edges are 25 M
    pow calls)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = len(lo)
    if binspacing == 'lin':
        start, stop, fn = np.minimum(0, lo), hi, np.linspace
    else:
        start, stop, fn = np.log10(np.maximum(lo, min_log_bins)), np.log10(hi), np.logspace
    if n * numbins < 1 << 20:
        return np.ascontiguousarray(fn(start, stop, numbins, axis=1))
    workers = max(1, min(16, os.cpu_count() or 1))
    step = -(-n // (4 * workers))
    blocks = [slice(i, min(n, i + step)) for i in range(0, n, step)]
    out = np.empty((n, numbins))

    def rows(sl):
        if binspacing == 'lin':
            out[sl] = np.linspace(start[sl], stop[sl], numbins, axis=1)
        else:       # np.logspace(start, stop, num) is np.power(10.0, np.linspace(start, stop, num))
            np.power(10.0, np.linspace(start[sl], stop[sl], numbins, axis=1), out=out[sl])

    with ThreadPoolExecutor(workers) as pool:
        list(pool.map(rows, blocks))
    return out


def integrate_batch(integrands, num_samples, *, edges=None, numbins=None, binspacing='log', min_log_bins=None,
                    seed=0, precision='f32', last_bc=None, first_bc=None, group=None):
    """A grid of parameter points in one launch per pass (BASELINE config 5).

    ``integrands``: list of Integrand sharing kind / coupling / observable accuracy (e.g. one ``theta_scale`` per
    theta_crit of gen_crit_sub_num_rad).  Every point sees the same sample stream.  ``edges``: array
    [n_points, n_edges]; or ``numbins`` + ``binspacing`` (+ per-point ``min_log_bins``) for the reference's
    data-dependent bins, found by a batched min/max pass.  Returns a dict of arrays with a leading n_points axis:
    bins, density, densityErr, integral, integralErr (same conventions as ``integrator``).

    With ``torch.distributed`` initialised the GRID is sharded, not the samples: every rank integrates its own
    contiguous block of parameter points over the whole sample stream, so nothing has to be reduced -- the ranks'
    finished rows are gathered ("shard the parameter grid instead of samples")."""
    from .montecarlo.integrator import _bc_args
    D.require_cuda()
    npts_all = len(integrands)
    world = torch.distributed.get_world_size(group) if dist_utils.is_distributed(group) else 1
    rank = torch.distributed.get_rank(group) if world > 1 else 0
    p0, p_count = dist_utils.shard_bounds(npts_all, rank, world)
    mine = list(integrands[p0:p0 + p_count])
    npts = len(mine)
    num_samples = int(num_samples)
    prec, key = PRECISIONS[precision], int(seed) & 0xFFFFFFFFFFFFFFFF
    arr = None                                      # the points' C structs, built once for both passes
    if edges is not None:
        edges = np.ascontiguousarray(edges, dtype=np.float64)[p0:p0 + p_count]
    elif npts:
        arr = _as_struct_array(mine)
        mm = D.to_device(np.tile(np.array([np.inf, -np.inf]), (npts, 1)))      # per point: [min, max] to be merged
        _lib.check(_lib.lib.jmc_run_batch(arr, npts, num_samples, key, 0, prec, None, 0, None, None, None, D.ptr(mm),
                                          D.stream()))
        mm = D.to_host(mm)
        mlb = np.full(npts_all, 1e-15) if min_log_bins is None else np.asarray(min_log_bins, dtype=float)
        edges = _set_bins_rows(mm[:, 0], mm[:, 1], mlb[p0:p0 + npts], int(numbins), binspacing)
    ne = edges.shape[1] if edges is not None and edges.ndim == 2 and edges.shape[0] else int(numbins)
    nb = ne - 1
    bc_kind, bc_value = _bc_args(first_bc, None if last_bc is None else list(last_bc))
    dens, err = D.empty((npts, nb)), D.empty((npts, nb))
    integ, ierr = D.empty((npts, ne)), D.empty((npts, nb))
    count = D.zeros((npts, nb), dtype=torch.int64)
    edges_dev = D.to_device(edges if npts else np.zeros((0, ne)))
    if npts:
        arr = arr if arr is not None else _as_struct_array(mine)
        sum_w, sum_w2 = D.zeros((npts, nb)), D.zeros((npts, nb))
        edges = np.ascontiguousarray(edges, dtype=np.float64)
        _lib.check(_lib.lib.jmc_run_batch(arr, npts, num_samples, key, 0, prec, edges.ctypes.data_as(C.c_void_p), ne,
                                          D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), None, D.stream()))
        areas = D.to_device(np.array([ig.area for ig in mine], dtype=np.float64))
        _lib.check(_lib.lib.jmc_finalize_batch(npts, D.ptr(sum_w), D.ptr(sum_w2), D.ptr(count), D.ptr(edges_dev), ne,
                                               D.ptr(areas), num_samples, bc_kind, bc_value, D.ptr(dens), D.ptr(err),
                                               D.ptr(integ), D.ptr(ierr), D.stream()))
    out = dict(bins=edges_dev.reshape(npts, ne), density=dens, densityErr=err, integral=integ, integralErr=ierr,
               count=count)
    if world > 1:
        out = {k: dist_utils.gather_rows(v, npts_all, world, group) for k, v in out.items()}
        return {k: D.to_host(v) for k, v in out.items()}
    # (one rank: the edges are on the host already)
    host = {k: D.to_host(v) for k, v in out.items() if k != 'bins'}
    host['bins'] = np.asarray(edges, dtype=np.float64).reshape(npts, ne)
    return host
