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
    mean = np.matmul(post_samples, w[:, :, None])[:, :, 0] / tw[:, None]
    zs = (post_samples - mean[:, :, None]) * np.sqrt(w)[:, None, :]
    cov = np.matmul(zs, np.transpose(zs, (0, 2, 1))) / (tw - 1.0)[:, None, None]
    return mean, cov


def bucket_empirical_stats(samples, rows_per_bucket):
    """Same statistics straight from the bucketed joint samples (no padded tensor): samples (nsample,T),
    rows_per_bucket = list of row-index arrays, one per kept maximizer."""
    T = samples.shape[1]
    S = len(rows_per_bucket)
    mean = np.empty((S, T))
    cov = np.empty((S, T, T))
    for j, rows in enumerate(rows_per_bucket):
        x = samples[rows]
        m = x.sum(axis=0) / len(rows)
        xc = x - m
        mean[j] = m
        cov[j] = xc.T.dot(xc) / (len(rows) - 1.0)
    return mean, cov
