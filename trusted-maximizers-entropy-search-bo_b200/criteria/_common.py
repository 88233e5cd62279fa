import numpy as np
import torch

from .. import ops
from ..device import dev


def to_np(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


def get_pNK_test_obs(ls, sigmas, sigma0s, nhyp, X, test_xs, dtype=None):
    """criteria/evaluate_mp.py:12-53 (identical copies in the other four criterion files).
    Returns (pNKs, invpNKs), each (nhyp, ntest+nobs, ntest+nobs).  The inverse is Cholesky-based while the matrix
    is numerically positive definite; the test block of pNK carries no noise or jitter, so near-duplicate
    test points can break that -- those matrices go through the general Gauss-Jordan inverse, which is what the
    reference's tf.linalg.inv (LU) computes for every matrix."""
    ls_, sg, s0 = to_np(ls), to_np(sigmas).reshape(-1), to_np(sigma0s).reshape(-1)
    ls_ = ls_.reshape(nhyp, -1)
    pNKs = torch.stack([ops.pnk(test_xs, X, ls_[i], float(sg[i]), float(s0[i])) for i in range(nhyp)])
    inv, info = ops.batched_chol2inv(pNKs, return_info=True)
    bad = torch.nonzero(info != 0).reshape(-1)
    if bad.numel():
        inv[bad] = ops.gj_inverse(pNKs[bad])
    if isinstance(X, np.ndarray):
        return pNKs.cpu().numpy(), inv.cpu().numpy()
    return pNKs, inv


def hyp(ls, sigmas, sigma0s, i, xdim):
    return to_np(ls).reshape(-1, xdim)[i], float(to_np(sigmas).reshape(-1)[i]), float(to_np(sigma0s).reshape(-1)[i])


def gather_nonzero(max_probs, *arrs):
    """tf.where(max_probs > 0) + tf.gather (evaluate_mp.py:169-173)."""
    p = dev(max_probs).reshape(-1)
    idx = torch.nonzero(p > 0.0).reshape(-1)
    return (p[idx],) + tuple(dev(a, p.device, dtype=(torch.uint8 if getattr(a, "dtype", None) in (np.bool_, torch.bool) else torch.float64))[idx] for a in arrs)


def finish(acc, like):
    return acc.cpu().numpy() if isinstance(like, np.ndarray) else acc
