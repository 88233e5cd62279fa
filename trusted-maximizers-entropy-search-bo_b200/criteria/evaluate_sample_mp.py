"""TES_sp, q = 1 -- mirror of criteria/evaluate_sample_mp.py."""
import torch

from .. import ops
from ._common import dev, finish, gather_nonzero, get_pNK_test_obs, hyp  # noqa: F401


def get_queried_f_stat_given_test_samples(x, l, sigma, sigma0, ntest, nobs, X, Y, test_xs, invpNK_test,
                                          invpNK_obs, post_test_samples, dtype=None):
    """evaluate_sample_mp.py:54-129 -> (query_mean (nmax,nx,nsample), query_var (nx,))."""
    inv = torch.cat([dev(invpNK_test), dev(invpNK_obs)], dim=1)
    M, b, Kq = ops.posterior_stats(x, test_xs, X, Y, l, float(sigma), inv)
    P = dev(post_test_samples)                                   # (nmax, ntest, K)
    nmax, T, K = P.shape
    qm = ops.gemm(M, P.permute(1, 0, 2).reshape(T, nmax * K)).reshape(-1, nmax, K).permute(1, 0, 2)
    qm = qm + b.reshape(1, -1, 1)
    return finish(qm.contiguous(), x), finish(Kq, x)


def mp(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs, max_probs_all,
       post_test_samples_all, post_test_mask_all, invpNK_all, dtype=None, niteration=10, use_loop=True,
       parallel_iterations=1, *, z=None, seed=0, n_global=None, x_offset=0):
    """evaluate_sample_mp.py:132-246 -> (nx,).  `z`: (nhyp, niteration, nysample, nmax', nx, nsample)."""
    avg = None
    masks = dev(post_test_mask_all, dtype=torch.uint8)
    for i in range(nhyp):
        l, sigma, sigma0 = hyp(ls, sigmas, sigma0s, i, xdim)
        p = dev(max_probs_all)[i].reshape(-1)
        idx = torch.nonzero(p > 0.0).reshape(-1)
        val = ops.sp_acq(x, test_xs, X, Y, l, sigma, sigma0, dev(invpNK_all)[i], p[idx],
                         dev(post_test_samples_all)[i][idx], masks[i][idx], niter=niteration, nysample=nysample,
                         z=None if z is None else z[i], seed=seed + i, n_global=n_global, x_offset=x_offset)
        avg = val / nhyp if avg is None else avg + val / nhyp
    return finish(avg, x)
