"""TES_sp, batch q -- mirror of criteria/evaluate_batch_sample_mp.py (x is (nx, batchsize, xdim))."""
import torch

from .. import ops
from ._common import dev, finish, get_pNK_test_obs, hyp  # noqa: F401


def mp(x, ls, sigmas, sigma0s, X, Y, xdim, nobs, nhyp, nysample, test_xs, max_probs_all, post_test_samples_all,
       post_test_mask_all, invpNK_all, dtype=None, niteration=10, use_loop=True, parallel_iterations=1, *, z=None,
       seed=0, n_global=None, x_offset=0):
    """evaluate_batch_sample_mp.py:160-290 -> (nx,).
    `z`: (nhyp, niteration, nysample, nmax', nx, nsample, batchsize)."""
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
