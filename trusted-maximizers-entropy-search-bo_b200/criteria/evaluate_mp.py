"""TES_ep, q = 1 -- mirror of criteria/evaluate_mp.py."""
import torch

from .. import ops
from ._common import dev, finish, gather_nonzero, get_pNK_test_obs, hyp  # noqa: F401


def get_queried_f_stat_given_data_maxidx(xs, l, sigma, sigma0, ntest, nobs, X, Y, test_xs, invpNK,
                                         post_mean_test_fs, post_cov_test_fs, dtype=None):
    """evaluate_mp.py:56-113 -> (query_mean, query_var), each (nmax, nx)."""
    p = torch.ones(dev(post_mean_test_fs).shape[0], dtype=torch.float64, device="cuda")
    mean, var = ops.ep_acq(ops.EP_MOMENTS, xs, test_xs, X, Y, l, float(sigma), float(sigma0), invpNK, None, p,
                           post_mean_test_fs, post_cov_test_fs)
    return finish(mean.t().contiguous(), xs), finish(var.t().contiguous(), xs)


def mp(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs, max_probs_all,
       post_mean_test_fs_all, post_cov_test_fs_all, invKobs_all, invpNK_all, stochastic=False, dtype=None,
       niteration=10, use_loop=True, parallel_iterations=1, *, z=None, seed=0, n_global=None, x_offset=0):
    """evaluate_mp.py:116-244 -> (nx,).  Keyword-only extras: `z` (nhyp, niteration, nysample, nmax', nx)
    host-supplied draws for the stochastic estimator, else Philox(`seed` + hyp index)."""
    avg = None
    for i in range(nhyp):
        l, sigma, sigma0 = hyp(ls, sigmas, sigma0s, i, xdim)
        p, pm, pc = gather_nonzero(dev(max_probs_all)[i], dev(post_mean_test_fs_all)[i],
                                   dev(post_cov_test_fs_all)[i])
        val = ops.ep_acq(ops.EP_STOCHASTIC if stochastic else ops.EP_DETERMINISTIC, x, test_xs, X, Y, l, sigma,
                         sigma0, dev(invpNK_all)[i], dev(invKobs_all)[i], p, pm, pc, niter=niteration,
                         nysample=nysample, z=None if z is None else z[i], seed=seed + i, n_global=n_global,
                         x_offset=x_offset)
        avg = val / nhyp if avg is None else avg + val / nhyp
    return finish(avg, x)
