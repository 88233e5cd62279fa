"""TES_ep, batch q -- mirror of criteria/evaluate_batch_mp.py.

As shipped, the reference scores its samples under the STANDARD normal (:302) and builds the marginal
from that same log-prob (:317-322), so the criterion equals -(sum p) log(sum p) ~ 0 for every input
(SURVEY Q5).  This mirror returns exactly that value (NaN for a query with a non-finite coordinate, as the
reference's kernel -> eigh -> log_prob chain does); it does not "fix" the estimator."""
from .. import ops
from ._common import dev, finish, gather_nonzero, get_pNK_test_obs  # noqa: F401


def mp(x, ls, sigmas, sigma0s, X, Y, xdim, nobs, nhyp, nysample, test_xs, max_probs_all, post_mean_test_fs_all,
       post_cov_test_fs_all, invKobs_all, invpNK_all, dtype=None, niteration=10, use_loop=True,
       parallel_iterations=1):
    """evaluate_batch_mp.py:145-267 -> (nx,)."""
    xt = dev(x)
    avg = None
    for i in range(nhyp):
        (p,) = gather_nonzero(dev(max_probs_all)[i])
        val = ops.batch_ep_acq(p, xt)
        avg = val / nhyp if avg is None else avg + val / nhyp
    return finish(avg, x)
