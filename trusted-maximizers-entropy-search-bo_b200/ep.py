"""Mirror of ep.approximate_EP_np (ep.py:73-204): EP for N(mu, Sigma) truncated to
{f_max_idx >= f_i for all i}; one site is updated per outer iteration and the Gaussian approximation is
refreshed by a full inverse (ep.py:182-185).

`approximate_EP_np` runs on the device (tes_ep_moments_f64; `approximate_EP_all` does every maximizer in one
launch, which is what utils.get_testidxs_stats(mode='ep') uses); `approximate_EP_host` is the host numpy/scipy
restatement, where the reference runs it."""
import numpy as np
import scipy.stats as spst


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


def _approx(cs, mean_xs, inv_cov_xs, site_means, site_vars):
    csp = cs / np.sqrt(site_vars.reshape(1, -1))
    cov = np.linalg.inv(inv_cov_xs + csp.dot(csp.T))
    mean = cov.dot(inv_cov_xs.dot(mean_xs)
                   + np.sum(cs * site_means.reshape(1, -1) / site_vars.reshape(1, -1), axis=1, keepdims=True))
    return mean, cov


def approximate_EP_host(max_idx, mean_xs, cov_xs, resolution=1e-9, max_niter=1000):
    with np.errstate(all="ignore"):
        mean_xs = np.asarray(mean_xs, dtype=np.float64).reshape(-1, 1)
        nx = mean_xs.shape[0]
        cs = np.zeros([nx, nx - 1])
        others = [i for i in range(nx) if i != max_idx]
        cs[max_idx, :] = 1.0
        cs[others, np.arange(nx - 1)] = -1.0
        site_means = cs.T.dot(mean_xs).reshape(nx - 1)
        site_vars = np.einsum("ij,ik,kj->j", cs, cov_xs, cs).copy()
        inv_cov = np.linalg.inv(cov_xs)
        a_mean, a_cov = _approx(cs, mean_xs, inv_cov, site_means, site_vars)
        count, diff = 0, 1e10
        while diff > resolution and count < max_niter:
            count += 1
            ctSc = np.einsum("ij,ik,kj->j", cs, a_cov, cs)
            cav_v = 1.0 / (1.0 / ctSc - 1.0 / site_vars)
            cav_m = cav_v * (cs.T.dot(a_mean).squeeze() / ctSc - site_means / site_vars)
            beta = cav_m / np.sqrt(cav_v)
            poc = spst.norm.pdf(beta) / spst.norm.cdf(beta)
            cf_m = cav_m + np.sqrt(cav_v) * poc
            cf_v = -cav_m * np.sqrt(cav_v) * poc - cf_m * cf_m + 2.0 * cav_m * cf_m - cav_m * cav_m + cav_v
            new_v = 1.0 / (1.0 / cf_v - 1.0 / cav_v)
            new_m = new_v * (cf_m / cf_v - cav_m / cav_v)
            uidx = count % (nx - 1)
            last = count - 1 if count >= 1 else nx - 2
            while (new_v[uidx] <= 0 or np.isnan(new_v[uidx]) or np.isinf(new_v[uidx])) and uidx != last:
                count += 1
                uidx = count % (nx - 1)
            site_means[uidx] = new_m[uidx]
            site_vars[uidx] = new_v[uidx]
            n_mean, n_cov = _approx(cs, mean_xs, inv_cov, site_means, site_vars)
            diff = np.mean((n_mean - a_mean) ** 2 + (n_cov - a_cov) ** 2)
            if np.any(np.isnan(n_cov)) or np.any(np.isnan(n_mean)):
                return a_mean, a_cov
            a_mean, a_cov = n_mean, n_cov
    return a_mean, a_cov
