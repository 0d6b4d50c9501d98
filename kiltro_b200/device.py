"""Device-buffer plumbing.  PyTorch is used ONLY to own device memory and name the current stream; every computation
goes through libkiltro_b200.so (kiltro_b200/_lib.py).  Nothing here computes on the CPU on behalf of the hot path."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.KiltroB200Error("kiltro_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Raw device address of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def empty(shape, dtype=torch.float64):
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def zeros(shape, dtype=torch.float64):
    return torch.zeros(shape, dtype=dtype, device=require_cuda())


def to_device(a, dtype):
    """NumPy array / torch tensor -> contiguous device tensor of `dtype` (host->device copy if needed)."""
    dev = require_cuda()
    if not isinstance(a, torch.Tensor):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dtype == dtype or a.is_cuda and not (a.dtype == torch.int64 and dtype == torch.int32):
        return a.to(device=dev, dtype=dtype).contiguous()
    if a.dtype == torch.int64 and dtype == torch.int32:
        # raw upload, then narrow on device with our own kernel (no host-side conversion pass)
        raw = a.to(device=dev).contiguous()
        out = torch.empty(raw.shape, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().kb_cast_i64_i32(raw.data_ptr(), raw.numel(), out.data_ptr(), stream_ptr()))
        return out
    return a.to(dtype).to(device=dev).contiguous()


def to_numpy(a):
    """Device tensor (or anything array-like) -> NumPy array on the host."""
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if hasattr(a, "to_numpy"):
        return a.to_numpy()
    return np.asarray(a)
