"""Mirror of the discrete-domain maximizer sampler of bo_discrete.py:359-400: an exact joint posterior sample
on the whole candidate grid, f = sqrtm(Sigma) z^T + mu with utils.sqrtm (SVD, utils.py:49-53), arg-max per
sample.  O(N^3): only viable for small grids (N ~ 400); the RFF path (stage A) is what scales.

Everything runs on the device: the posterior mean / covariance (utils.compute_mean_var_f, fullcov=True), the
matrix square root (batched Jacobi eigen-solver: for a symmetric matrix V E V^T the SVD factors are U = V sign(E),
S = |E|, so U sqrt(S) V^T = V sign(E) sqrt|E| V^T), the colouring GEMM and the arg-max.  The rounding-level part of
the spectrum (|e| ~ 1e-17 |cov|, the null space of the posterior covariance) enters sqrtm at ~3e-9 with LAPACK-
dependent singular vectors in the reference too; function values agree to ~1e-8, arg-max indices exactly."""
import numpy as np
import torch

from . import ops, utils
from .device import dev


def sqrtm(mat):
    """utils.sqrtm (utils.py:49-53): U diag(sqrt(s)) V^T of an SVD, for the symmetric matrices it is called on."""
    A = dev(mat)
    e, v = ops.sym_eig(A)
    out = ops.gemm(v * (torch.sign(e) * torch.sqrt(e.abs())).reshape(1, -1), v, transb=True)
    return out if torch.is_tensor(mat) else out.cpu().numpy()


def sample_maximizers(xs, X, Y, l, sigma, sigma0, z):
    """z: (nmax, nx) standard normals (bo_discrete.py:359-362).  Returns (sample_max_xs (nmax,d),
    sample_max_fs (nmax,), max_idx (nmax,), mean (nx,), cov (nx,nx))."""
    xs_d = dev(xs)
    mean, cov = utils.compute_mean_var_f(xs_d, dev(X), dev(Y), l, sigma, sigma0, fullcov=True)
    sq = sqrtm(cov)
    f = ops.gemm(dev(z), sq, transb=True) + mean.reshape(1, -1)     # (nmax, nx) = z sqrt^T + mu
    idx = torch.stack([ops.argmax(f[i])[0][0] for i in range(f.shape[0])])
    fs = f.gather(1, idx.reshape(-1, 1)).reshape(-1)
    out = (xs_d[idx], fs, idx, mean, cov)
    if isinstance(xs, np.ndarray):
        return tuple(o.cpu().numpy() for o in out)
    return out
