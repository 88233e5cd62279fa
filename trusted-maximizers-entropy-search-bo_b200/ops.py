"""Thin typed wrappers over the C ABI (include/tes_b200.h) for CUDA torch tensors."""
import torch

from . import _lib
from .device import WORKSPACE, dev, ptr, stream_ptr


def _rff_args(cand, W, b, theta):
    cand = dev(cand)
    N, d = cand.shape
    theta = dev(theta)
    theta = theta.reshape(theta.shape[0], -1)
    S, F = theta.shape
    W = dev(W).reshape(-1, F, d)
    b = dev(b).reshape(-1, F)
    if W.shape[0] not in (1, S) or b.shape[0] != W.shape[0]:
        raise ValueError("W must be (S,F,d) or (1,F,d) and b (S,F[,1]) or (1,F[,1]) to match theta (S,F[,1])")
    w_shared = int(W.shape[0] == 1 and S > 1)
    return cand, W, b, theta, N, d, S, F, w_shared


def rff_topk(cand, W, b, theta, sigma, k=1, idx_offset=0):
    """Top-k candidates of every RFF function sample -> (idx (S,k) int64, val (S,k)) CUDA tensors.
    utils_for_continuous.py:172-187."""
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    L = _lib.lib()
    need = L.tes_rff_topk_workspace_bytes(N, d, S, F, k)
    if need < 0:
        raise ValueError("invalid sizes for tes_rff_topk_f64: N=%d d=%d S=%d F=%d k=%d" % (N, d, S, F, k))
    ws = WORKSPACE.get(need, cand.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=cand.device)
    out_val = torch.empty((S, k), dtype=torch.float64, device=cand.device)
    rc = L.tes_rff_topk_f64(ptr(cand), N, d, ptr(W), ptr(b), ptr(theta), S, F, ws_, float(sigma), int(k),
                            int(idx_offset), ptr(out_idx), ptr(out_val), ptr(ws), ws.numel(), stream_ptr())
    _lib.check(rc, "tes_rff_topk_f64")
    return out_idx, out_val


def rff_eval(cand, W, b, theta, sigma):
    """All function-sample values (S, N).  optfunc.py:389-402 / utils_for_continuous.py:257-296."""
    cand, W, b, theta, N, d, S, F, ws_ = _rff_args(cand, W, b, theta)
    out = torch.empty((S, N), dtype=torch.float64, device=cand.device)
    rc = _lib.lib().tes_rff_eval_f64(ptr(cand), N, d, ptr(W), ptr(b), ptr(theta), S, F, ws_, float(sigma),
                                     ptr(out), stream_ptr())
    _lib.check(rc, "tes_rff_eval_f64")
    return out


def topk_merge(vals, idx):
    """(P,S,k) partial lists -> (S,k) under (value desc, index asc)."""
    vals = dev(vals)
    idx = dev(idx, dtype=torch.int64)
    P, S, k = vals.shape
    out_val = torch.empty((S, k), dtype=torch.float64, device=vals.device)
    out_idx = torch.empty((S, k), dtype=torch.int64, device=vals.device)
    rc = _lib.lib().tes_topk_merge_f64(ptr(vals), ptr(idx), P, S, k, ptr(out_val), ptr(out_idx), stream_ptr())
    _lib.check(rc, "tes_topk_merge_f64")
    return out_idx, out_val


def fp64_fma_peak(seconds=0.5, device=None):
    """Sustained DFMA rate (TFLOP/s) of the current device, timed with CUDA events."""
    import ctypes
    L = _lib.lib()
    dev_ = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    nsm = torch.cuda.get_device_properties(dev_).multi_processor_count
    blocks = nsm * 8
    sink = torch.empty(blocks * 256, dtype=torch.float64, device=dev_)
    flops = ctypes.c_double(0.0)
    iters = 1 << 16

    def run(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.tes_fp64_fma_burn(blocks, it, ptr(sink), ctypes.byref(flops), stream_ptr()), "burn")
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3, flops.value

    run(iters)
    t, fl = run(iters)
    it2 = max(iters, int(iters * seconds / max(t, 1e-6)))
    t, fl = run(it2)
    return fl / t * 1e-12
