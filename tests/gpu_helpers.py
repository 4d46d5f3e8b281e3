"""Thin test-side wrappers that call the C ABI of libjmc_b200.so directly
(device pointers from torch tensors), so the parity tests exercise exactly the
entry points include/jmc_b200.h declares."""
import ctypes as C

import numpy as np
import torch

from jetmontecarlo_b200 import _device as D
from jetmontecarlo_b200 import _lib

L = _lib.lib


def dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype).contiguous()


def uniforms(n, seed, first, n_uniform):
    out = D.empty((n, n_uniform))
    _lib.check(L.jmc_uniforms(n, seed, first, n_uniform, D.ptr(out), D.stream()))
    return out.cpu().numpy()


def sample(kind, method, n, seed, first=0, epsilon=0., zc=0., radius=1., lo=0., hi=1.):
    s = _lib.Sampler(kind=kind, method=_lib.METHODS[method], epsilon=epsilon, zc=zc, radius=radius, lo=lo, hi=hi)
    dim = 2 if kind in (_lib.SAMPLER_CRITICAL, _lib.SAMPLER_UNGROOMED) else 1
    samples = D.empty((n, 2) if dim == 2 else (n,))
    jac = D.empty((n,))
    _lib.check(L.jmc_sample(C.byref(s), n, seed, first, D.ptr(samples), D.ptr(jac), D.stream()))
    return samples.cpu().numpy(), jac.cpu().numpy()


def eval_integrand(g, n, crit=None, sub=None, pre=None, zb_u=None, edges=None):
    """returns jac, obs, weight*jac, bin (bin None without edges)"""
    s = g.c_struct()
    t = {k: (None if v is None else dev(v)) for k, v in dict(crit=crit, sub=sub, pre=pre, zb=zb_u, edges=edges).items()}
    jac, obs, w = D.empty((n,)), D.empty((n,)), D.empty((n,))
    bins = None if edges is None else D.empty((n,), dtype=torch.int32)
    _lib.check(L.jmc_eval_integrand(C.byref(s), n, D.ptr(t['crit']), D.ptr(t['sub']), D.ptr(t['pre']), D.ptr(t['zb']),
                                    D.ptr(t['edges']), 0 if edges is None else len(edges), D.ptr(jac), D.ptr(obs),
                                    D.ptr(w), D.ptr(bins), D.stream()))
    return (jac.cpu().numpy(), obs.cpu().numpy(), w.cpu().numpy(), None if bins is None else bins.cpu().numpy())


def bin_index(obs, edges):
    o, e = dev(obs), dev(edges)
    out = D.empty((len(obs),), dtype=torch.int32)
    _lib.check(L.jmc_bin_index(D.ptr(o), len(obs), D.ptr(e), len(edges), D.ptr(out), D.stream()))
    return out.cpu().numpy()


def hist(obs, w, edges):
    o, e = dev(obs), dev(edges)
    wd = None if w is None else dev(w)
    nb = len(edges) - 1
    sw, sw2, cnt = D.zeros((nb,)), D.zeros((nb,)), D.zeros((nb,), dtype=torch.int64)
    _lib.check(L.jmc_hist(D.ptr(o), D.ptr(wd), len(obs), D.ptr(e), len(edges), D.ptr(sw), D.ptr(sw2), D.ptr(cnt),
                          D.stream()))
    return sw.cpu().numpy(), sw2.cpu().numpy(), cnt.cpu().numpy()


def run(g, n, seed, edges, precision='f64', first=0, want_minmax=False, accumulate_into=None):
    s = g.c_struct()
    e = None if edges is None else dev(edges)
    nb = 0 if edges is None else len(edges) - 1
    if accumulate_into is None:
        sw, sw2, cnt = D.zeros((nb,)), D.zeros((nb,)), D.zeros((nb,), dtype=torch.int64)
    else:
        sw, sw2, cnt = accumulate_into
    mm = torch.tensor([np.inf, -np.inf], dtype=torch.float64, device="cuda") if want_minmax else None
    _lib.check(L.jmc_run(C.byref(s), n, seed, first, {'f64': 0, 'f32': 1}[precision], D.ptr(e),
                         0 if edges is None else len(edges), D.ptr(sw if nb else None), D.ptr(sw2 if nb else None),
                         D.ptr(cnt if nb else None), D.ptr(mm), D.stream()))
    torch.cuda.synchronize()
    res = dict(sum_w=sw, sum_w2=sw2, count=cnt)
    if want_minmax:
        res['minmax'] = mm.cpu().numpy()
    return res


def finalize(sw, sw2, cnt, edges, area, n_total, bc_kind, bc_value):
    e = dev(edges)
    nb = len(edges) - 1
    d, de, integ, ie = D.empty((nb,)), D.empty((nb,)), D.empty((nb + 1,)), D.empty((nb,))
    _lib.check(L.jmc_finalize(D.ptr(sw), D.ptr(sw2), D.ptr(cnt), D.ptr(e), nb + 1, float(area), int(n_total),
                              bc_kind, float(bc_value), D.ptr(d), D.ptr(de), D.ptr(integ), D.ptr(ie), D.stream()))
    return d.cpu().numpy(), de.cpu().numpy(), integ.cpu().numpy(), ie.cpu().numpy()


def dump(g, n, seed, precision='f64', first=0):
    """per-sample view of the fused streams: samples [n,8], obs [n], weight*jac [n]"""
    s = g.c_struct()
    samples, obs, w = D.empty((n, 8)), D.empty((n,)), D.empty((n,))
    _lib.check(L.jmc_dump(C.byref(s), n, seed, first, {'f64': 0, 'f32': 1}[precision], D.ptr(samples), D.ptr(obs),
                          D.ptr(w), D.stream()))
    return samples.cpu().numpy(), obs.cpu().numpy(), w.cpu().numpy()
