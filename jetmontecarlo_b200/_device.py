"""Device plumbing: PyTorch owns device memory and streams, nothing else.

Every kernel of the path is in libjmc_b200.so; torch is used only to allocate
tensors, move bytes between host and device and (in distributed.py) for the
NCCL all-reduce of the histograms.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_F64 = torch.float64


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("jetmontecarlo_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback for this path.")


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    """cudaStream_t of torch's current stream, as the void* the C ABI takes."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def is_device_tensor(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def to_device(x):
    """Host array / tensor -> contiguous float64 CUDA tensor (no copy if already one)."""
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.to(device())
        if x.dtype == _F64 and x.dim() == 1 and not x.is_contiguous() and x.stride(0) > 0:
            return _gather_strided(x)
        return x.to(_F64).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.nbytes >= 2 * _STAGE_BYTES:
        return _upload_staged(a)
    return torch.from_numpy(a).to(device(), non_blocking=False)


def _staging(dev):
    if dev not in _stage:
        _stage[dev] = ([torch.empty(_STAGE_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                       torch.cuda.Stream(device=torch.device("cuda", dev)), [torch.cuda.Event() for _ in range(2)])
    return _stage[dev]


def _upload_staged(a):
    """Large host array -> device tensor through the two pinned staging buffers: the host copy of chunk k + 1 into
    its buffer (four threads) overlaps the host-to-device copy of chunk k (the reference's API hands over host
    arrays, so for large N these copies are what its drop-in calls cost)."""
    dev = device()
    bufs, side, events = _staging(dev.index)
    out = torch.empty(a.shape, dtype=_F64, device=dev)
    flat_out = out.reshape(-1).view(torch.uint8)
    flat_in = a.reshape(-1).view(np.uint8)
    total = flat_in.size
    side.wait_stream(torch.cuda.current_stream(dev))
    for k, o in enumerate(range(0, total, _STAGE_BYTES)):
        nbytes = min(_STAGE_BYTES, total - o)
        events[k & 1].synchronize()                 # the copy that last read this buffer has finished
        dst = bufs[k & 1][:nbytes].numpy()
        quarter = -(-nbytes // 4)
        list(_copy_pool().map(lambda j: dst[j:j + quarter].__setitem__(slice(None), flat_in[o + j:o + min(nbytes, j + quarter)]),
                              range(0, nbytes, quarter)))
        with torch.cuda.stream(side):
            flat_out[o:o + nbytes].copy_(bufs[k & 1][:nbytes], non_blocking=True)
            events[k & 1].record(side)
    torch.cuda.current_stream(dev).wait_stream(side)
    return out


def _gather_strided(x):
    """contiguous copy of a strided 1-D FP64 device view (a column of an (N, 2) sample array) by the library's own
    strided-read kernel: x * 1.0 through jmc_eval_fn, which takes an element stride per argument"""
    out = torch.empty((x.numel(),), dtype=_F64, device=x.device)
    params = _lib.FnParams(s1=1.)
    null, one = C.c_void_p(0), C.c_int64(1)
    _lib.check(_lib.lib.jmc_eval_fn(_lib.FN_MUL, C.byref(params), x.numel(), ptr(x), C.c_int64(x.stride(0)), null, one,
                                    null, one, null, one, ptr(out), stream()))
    return out


def empty(shape, dtype=_F64):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=_F64):
    """zero-filled device tensor: allocation by torch, the fill by cudaMemsetAsync on torch's stream (jmc_memset)"""
    t = torch.empty(shape, dtype=dtype, device=device())
    zero_(t)
    return t


def zero_(t):
    """t[...] = 0 for a contiguous device tensor"""
    assert t.is_contiguous()
    _lib.check(_lib.lib.jmc_memset(ptr(t), 0, t.numel() * t.element_size(), stream()))
    return t


def full(shape, value):
    """FP64 device tensor filled with ``value`` (jmc_fill)"""
    t = torch.empty(shape, dtype=_F64, device=device())
    _lib.check(_lib.lib.jmc_fill(ptr(t), t.numel(), float(value), stream()))
    return t


_STAGE_BYTES = 32 << 20        # per pinned staging buffer
_stage = {}                     # device index -> (two pinned uint8 buffers, copy stream, two events)


def to_host(t):
    """Device tensor -> NumPy array.  Large results (BASELINE config 5 returns ~1 GB of rows) go through two pinned
    staging buffers on a side stream: the device-to-host copy of chunk k+1 runs while chunk k is copied from the
    staging buffer into the result (a plain ``.cpu()`` into pageable memory runs at ~2 GB/s)."""
    if not t.is_cuda or t.numel() * t.element_size() < 2 * _STAGE_BYTES:
        return t.cpu().numpy()
    t = t.contiguous()
    bufs, side, events = _staging(t.device.index)
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    flat_out = out.reshape(-1).view(np.uint8)
    flat_in = t.reshape(-1).view(torch.uint8)
    total = flat_in.numel()
    side.wait_stream(torch.cuda.current_stream(t.device))
    chunks = [(o, min(_STAGE_BYTES, total - o)) for o in range(0, total, _STAGE_BYTES)]

    def issue(k):
        o, nbytes = chunks[k]
        with torch.cuda.stream(side):
            bufs[k & 1][:nbytes].copy_(flat_in[o:o + nbytes], non_blocking=True)
            events[k & 1].record(side)

    def unload(k):
        # staging buffer -> result, in four pieces on the host's cores (first touch of the result's pages is the
        # expensive part; NumPy's copy loops release the GIL)
        o, nbytes = chunks[k]
        src = bufs[k & 1][:nbytes].numpy()
        quarter = -(-nbytes // 4)
        list(_copy_pool().map(lambda j: flat_out[o + j:o + min(nbytes, j + quarter)].__setitem__(slice(None), src[j:j + quarter]),
                              range(0, nbytes, quarter)))

    issue(0)
    for k in range(len(chunks)):
        if k + 1 < len(chunks):
            issue(k + 1)
        events[k & 1].synchronize()
        unload(k)
        # (chunk k + 2 reuses this buffer: it is issued only after this host copy has finished)
    return out


_pool = None


def _copy_pool():
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(4)
    return _pool


def check(rc):
    return _lib.check(rc)
