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
        return x.to(_F64).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(a).to(device(), non_blocking=False)


def empty(shape, dtype=_F64):
    return torch.empty(shape, dtype=dtype, device=device())


def zeros(shape, dtype=_F64):
    return torch.zeros(shape, dtype=dtype, device=device())


def to_host(t):
    return t.cpu().numpy()


def check(rc):
    return _lib.check(rc)
