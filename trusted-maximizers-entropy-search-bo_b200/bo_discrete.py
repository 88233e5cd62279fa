"""Mirror of the discrete-domain maximizer sampler of bo_discrete.py:359-400: an exact joint posterior sample
on the whole candidate grid, f = sqrtm(Sigma) z^T + mu with utils.sqrtm (SVD, utils.py:49-53), arg-max per
sample.  O(N^3): only viable for small grids (N ~ 400); the RFF path (stage A) is what scales.

The posterior mean / covariance (utils.compute_mean_var_f, fullcov=True), the colouring GEMM and the arg-max
run on the device; the N x N SVD square root is taken on the host with numpy, whose U sqrt(S) V^T is the
reference's definition of sqrtm."""
import numpy as np
import torch

from . import ops, utils
from .device import dev


def sqrtm(mat):
    """utils.sqrtm (utils.py:49-53): U diag(sqrt(s)) V^T from an SVD."""
    m = mat.detach().cpu().numpy() if torch.is_tensor(mat) else np.asarray(mat)
    U, s, Vt = np.linalg.svd(m)
    return (U * np.sqrt(s)) @ Vt


def sample_maximizers(xs, X, Y, l, sigma, sigma0, z):
    """z: (nmax, nx) standard normals (bo_discrete.py:359-362).  Returns (sample_max_xs (nmax,d),
    sample_max_fs (nmax,), max_idx (nmax,), mean (nx,), cov (nx,nx))."""
    xs_d = dev(xs)
    mean, cov = utils.compute_mean_var_f(xs_d, dev(X), dev(Y), l, sigma, sigma0, fullcov=True)
    sq = dev(sqrtm(cov))
    f = ops.gemm(dev(z), sq, transb=True) + mean.reshape(1, -1)     # (nmax, nx) = z sqrt^T + mu
    idx = torch.stack([ops.argmax(f[i])[0][0] for i in range(f.shape[0])])
    fs = f.gather(1, idx.reshape(-1, 1)).reshape(-1)
    out = (xs_d[idx], fs, idx, mean, cov)
    if isinstance(xs, np.ndarray):
        return tuple(o.cpu().numpy() for o in out)
    return out
