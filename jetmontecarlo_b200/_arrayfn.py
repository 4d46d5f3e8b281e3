"""Array functions of the replay path: NumPy-style arguments in, one CUDA
kernel (jmc_eval_fn), NumPy-style result out.

Arguments may be Python scalars, NumPy arrays (any shape, broadcast against
each other) or CUDA tensors.  Host inputs give a host result of the broadcast
shape (a NumPy float64 scalar when every argument is a scalar); if any input
is a CUDA tensor the result stays on the device.
"""
import ctypes as C

import numpy as np
import torch

from . import _device as D
from . import _lib


def _is_scalar(x):
    return np.ndim(x) == 0 and not isinstance(x, torch.Tensor)


def eval_fn(fn, args, jet=_lib.QUARK, acc=_lib.ACC_LL, fixed_coupling=True, z_cut=0., beta=2.,
            epsilon=0., scalar_fields=None):
    """Run array function ``fn`` (an ``_lib.FN_*`` id).

    ``args``: up to four entries, each a scalar or array.
    ``scalar_fields``: {arg position: FnParams field name} for arguments whose
    scalar form travels in a named field (max_radius, z_pre, f_soft).
    """
    D.require_cuda()
    args = list(args) + [None] * (4 - len(args))
    on_device = any(D.is_device_tensor(a) for a in args)
    arrays = {}
    params = _lib.FnParams(jet=jet, acc=acc, fixed_coupling=int(bool(fixed_coupling)), z_cut=float(z_cut),
                           beta=float(beta), max_radius=1., z_pre=0., f_soft=1., epsilon=float(epsilon))
    scalar_fields = scalar_fields or {}
    for k, a in enumerate(args):
        if a is None:
            continue
        if _is_scalar(a):
            setattr(params, "s%d" % k, float(a))
            if k in scalar_fields:
                setattr(params, scalar_fields[k], float(a))
        else:
            arrays[k] = a
    if not arrays:
        shape = ()
        n = 1
    else:
        shapes = [tuple(a.shape) for a in arrays.values()]
        shape = np.broadcast_shapes(*shapes)
        n = int(np.prod(shape, dtype=np.int64))
    dev, stride_of = {}, {}
    for k, a in arrays.items():
        if (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.float64 and a.dim() == 1
                and tuple(a.shape) == shape and a.stride(0) >= 1):
            dev[k], stride_of[k] = a, a.stride(0)      # a column of an (N, 2) sample array is read in place
            continue
        if tuple(a.shape) != shape:
            a = torch.broadcast_to(a, shape) if isinstance(a, torch.Tensor) else np.broadcast_to(a, shape)
        dev[k] = D.to_device(a).reshape(-1)
    out = D.empty((n,))
    ptrs, strides = [], []
    for k in range(4):
        t = dev.get(k)
        ptrs.append(D.ptr(t))
        strides.append(C.c_int64(stride_of.get(k, 1)))
    _lib.check(_lib.lib.jmc_eval_fn(fn, C.byref(params), n, ptrs[0], strides[0], ptrs[1], strides[1],
                                    ptrs[2], strides[2], ptrs[3], strides[3], D.ptr(out), D.stream()))
    if on_device:
        return out.reshape(shape)
    res = D.to_host(out).reshape(shape)
    return res[()] if shape == () else res
