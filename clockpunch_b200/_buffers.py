"""Device-buffer helpers: PyTorch tensors are used ONLY as zero-copy CUDA buffers
(allocation, the current stream, host<->device copies); all arithmetic happens in
libpunch_b200 kernels reached through ctypes."""
from __future__ import annotations

import numpy as np

from . import _lib


def torch_mod():
    import torch
    return torch


def device_of(ctx: "_lib.Context"):
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise _lib.PunchError("no CUDA device available: clockpunch_b200 needs a GPU (no CPU fallback)")
    return torch.device("cuda", ctx.device)


def to_device(a: np.ndarray, ctx, dtype=np.float64):
    torch = torch_mod()
    a = np.ascontiguousarray(a, dtype=dtype)
    return torch.from_numpy(a).to(device_of(ctx), non_blocking=False)


def empty(shape, ctx, dtype=None):
    torch = torch_mod()
    return torch.empty(shape, dtype=dtype or torch.float64, device=device_of(ctx))


def stream_ptr(ctx) -> int:
    torch = torch_mod()
    return torch.cuda.current_stream(device_of(ctx)).cuda_stream


class _RawDeviceBytes:
    """__cuda_array_interface__ over memory the library owns (pc_comm_table's gathered buffer)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr: int, nbytes: int, device):
    """uint8 tensor aliasing `nbytes` of device memory at `ptr` (no copy, no ownership)."""
    torch = torch_mod()
    with torch.cuda.device(device):
        return torch.as_tensor(_RawDeviceBytes(ptr, nbytes), device=device)
