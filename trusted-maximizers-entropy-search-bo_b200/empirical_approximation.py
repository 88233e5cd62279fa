"""Mirror of empirical_approximation.get_empirical_stat_from_samples (empirical_approximation.py:86-107),
computed on the device (tes_bucket_moments_f64)."""
import numpy as np
import torch

from . import ops
from .device import dev


def get_empirical_stat_from_samples(samples, weight=None):
    """samples (K, T), weight (K,) -> mean (T, 1), cov (T, T) = sum_k w_k (x_k - m)(x_k - m)^T / (sum w - 1).
    The reference only ever passes the 0/1 validity mask of the padded sample tensor as `weight`
    (utils.py:1811-1817); other weights are rejected."""
    like = samples
    x = dev(samples)
    K, T = x.shape
    if weight is None:
        rows = torch.arange(K, dtype=torch.int64, device=x.device)
    else:
        w = dev(weight).reshape(-1)
        if w.numel() != K or bool(((w != 0) & (w != 1)).any()):
            raise ValueError("weight must be a 0/1 mask of length samples.shape[0]")
        rows = torch.nonzero(w != 0).reshape(-1).contiguous()
    starts = torch.tensor([0, rows.numel()], dtype=torch.int64, device=x.device)
    cols = torch.arange(T, dtype=torch.int64, device=x.device)
    mean, cov = ops.bucket_moments(x, rows, starts, cols)
    mean, cov = mean.reshape(T, 1), cov.reshape(T, T)
    if isinstance(like, np.ndarray):
        return mean.cpu().numpy(), cov.cpu().numpy()
    return mean, cov
