"""TES_ep closed-form "lite" -- mirror of criteria/evaluate_mp_lite.py (imported but never called by
the reference's drivers)."""
from .. import ops
from ._common import dev, finish, gather_nonzero, get_pNK_test_obs, hyp  # noqa: F401
from .evaluate_mp import get_queried_f_stat_given_data_maxidx  # noqa: F401


def mp(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, test_xs, max_probs_all, post_mean_test_fs_all,
       post_cov_test_fs_all, invpNK_all, dtype=None):
    """evaluate_mp_lite.py:116-225 -> (nx,)."""
    avg = None
    for i in range(nhyp):
        l, sigma, sigma0 = hyp(ls, sigmas, sigma0s, i, xdim)
        p, pm, pc = gather_nonzero(dev(max_probs_all)[i], dev(post_mean_test_fs_all)[i],
                                   dev(post_cov_test_fs_all)[i])
        val = ops.ep_acq(ops.EP_LITE, x, test_xs, X, Y, l, sigma, sigma0, dev(invpNK_all)[i], None, p, pm, pc)
        avg = val / nhyp if avg is None else avg + val / nhyp
    return finish(avg, x)
