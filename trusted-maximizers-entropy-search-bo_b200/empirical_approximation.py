"""Mirror of empirical_approximation.get_empirical_stat_from_samples (empirical_approximation.py:86-107)."""
import numpy as np


def get_empirical_stat_from_samples(samples, weight=None):
    samples = np.asarray(samples)
    n = samples.shape[0]
    if weight is None:
        weight = np.ones(n)
    tw = np.sum(weight)
    w = np.asarray(weight).reshape(-1, 1)
    emean = np.sum(samples * w, axis=0, keepdims=True) / tw
    zs = (samples - emean) * np.sqrt(w)
    return emean.reshape(-1, 1), zs.T.dot(zs) / (tw - 1.0)


def batched_empirical_stats(post_samples, post_masks):
    """All maximizers of one hyper-parameter sample at once: post_samples (S,T,K), post_masks (S,K)
    -> means (S,T), covs (S,T,T); identical to calling the function above per maximizer."""
    w = post_masks.astype(np.float64)                       # (S,K)
    tw = w.sum(axis=1)                                      # (S,)
    mean = np.einsum("stk,sk->st", post_samples, w) / tw[:, None]
    zs = (post_samples - mean[:, :, None]) * np.sqrt(w)[:, None, :]
    cov = np.einsum("sak,sbk->sab", zs, zs) / (tw - 1.0)[:, None, None]
    return mean, cov
