"""Device-memory plumbing (PyTorch is used only for allocation, streams and host<->device copies)."""
import numpy as np
import torch

from ._lib import TesError


def require_cuda():
    if not torch.cuda.is_available():
        raise TesError("no CUDA device: the TES kernels are sm_100a-only and have no CPU fallback")


def dev(x, device=None, dtype=torch.float64):
    """numpy / torch (CPU or CUDA) -> contiguous CUDA tensor of `dtype`."""
    require_cuda()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class Workspace:
    """Grow-only scratch buffer handed to the C ABI (no hidden allocation in the library), one per (device, CUDA
    stream): ops issued on different streams (or from different host threads on their own streams) never share
    scratch; ops on one stream are ordered by the stream."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


WORKSPACE = Workspace()
