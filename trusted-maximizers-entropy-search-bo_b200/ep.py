"""Mirror of ep.approximate_EP_np (ep.py:73-204): EP for N(mu, Sigma) truncated to
{f_max_idx >= f_i for all i}; one site is updated per outer iteration and the Gaussian approximation is
refreshed by a full inverse (ep.py:182-185).

`approximate_EP_np` runs on the device (tes_ep_moments_f64; `approximate_EP_all` does every maximizer in one
launch, which is what utils.get_testidxs_stats(mode='ep') uses).  There is no host implementation in this
package: the numpy/scipy restatement the tests check against is test infrastructure, outside this package."""
import numpy as np


def approximate_EP_all(mean_xs, cov_xs, resolution=1e-9, max_niter=1000):
    """All maximizers at once on the device: mean (T,) / (T,1), cov (T,T) -> post_mean (T,T), post_cov (T,T,T)
    (numpy in -> numpy out, CUDA tensors in -> CUDA tensors out)."""
    import torch
    from . import ops
    pm, pc = ops.ep_moments(mean_xs, cov_xs, resolution, max_niter)
    if torch.is_tensor(cov_xs):
        return pm[0], pc[0]
    return pm[0].cpu().numpy(), pc[0].cpu().numpy()


def approximate_EP_np(max_idx, mean_xs, cov_xs, resolution=1e-9, max_niter=1000):
    """ep.py:73-204 (the reference's signature): (T,1) mean and (T,T) covariance of the EP approximation for
    maximizer `max_idx`, computed on the device."""
    pm, pc = approximate_EP_all(np.asarray(mean_xs, dtype=np.float64), np.asarray(cov_xs, dtype=np.float64),
                                resolution, max_niter)
    return pm[max_idx].reshape(-1, 1), pc[max_idx]
