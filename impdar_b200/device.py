"""Buffer carrier: PyTorch tensors hold device memory and streams; nothing here computes.

A radargram travels as ``dat.data``: a host numpy array (the reference's contract,
RadarData/__init__.py:136-137) or - the device-resident lane - a CUDA ``torch.Tensor`` that then stays
on the GPU across calls (SURVEY.md 8b "optional fast lane").
"""
import ctypes

import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("impdar_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def is_device_array(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def current_stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a tensor / numpy array as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    if hasattr(t, "data_ptr"):                  # torch tensors and raw-address blocks (parallel._RawBlock)
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def to_device(x, dtype=torch.float32):
    """Contiguous CUDA tensor of ``dtype`` holding x (numpy array or tensor)."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()
    a = np.asarray(x)
    np_dtype = {torch.float32: np.float32, torch.float64: np.float64}[dtype]
    if a.dtype != np_dtype or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np_dtype)
    return torch.from_numpy(a).cuda(non_blocking=True)


def host_f64(x):
    """Small vector as contiguous host float64 numpy array (kept alive by the caller for the call)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


_workspaces = {}


def workspace(nbytes):
    """A cached scratch buffer of at least nbytes (uint8 CUDA tensor), one per (device, current stream): calls
    enqueued on different streams (impdar_b200.process keeps several profiles in flight) never share scratch."""
    require_cuda()
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = None
        ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device="cuda")
        _workspaces[key] = ws
    return ws


def free_workspaces():
    _workspaces.clear()


def to_host(t, np_dtype):
    """CUDA tensor -> host numpy array of np_dtype.

    The dtype conversion runs on the device and the copy lands in page-locked memory (torch's caching host
    allocator), so the transfer is one DMA at PCIe speed instead of a staged copy plus a single-core cast.
    The returned array owns that buffer (it is recycled when the array is garbage collected)."""
    import torch
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float64:
        t = t.double()
    elif np_dtype == np.float32:
        t = t.float()
    t = t.contiguous()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out = host.numpy()
    if out.dtype != np_dtype:
        out = out.astype(np_dtype)
    return out
