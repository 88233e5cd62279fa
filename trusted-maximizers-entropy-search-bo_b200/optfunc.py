"""Mirror of the reference's optfunc.py (numpy half): RFF posterior weight draw and function-sample
evaluation.  The draw is host numpy in the reference (bo_batch.py:641-652) and stays so; evaluation
runs on the device."""
import numpy as np

from . import ops


def draw_random_init_weights_features_np(xdim, n_funcs, n_features, xx, yy, l, sigma, sigma0, draws=None):
    """optfunc.py:303-385.  `draws` = (randn_W (S,F,d), rand_b (S,F,1), randn_noise (S,F,1)) makes the
    three global-RNG draws (:329, :334, :344) explicit; omitted -> np.random in the same order."""
    n = xx.shape[0]
    l = np.asarray(l, dtype=np.float64).reshape(1, xdim)
    if draws is None:
        randn_W = np.random.randn(n_funcs, n_features, xdim)
        rand_b = np.random.rand(n_funcs, n_features, 1)
    else:
        randn_W, rand_b = draws[0], draws[1]
    W = randn_W * np.sqrt(l).reshape(1, 1, xdim)
    b = 2.0 * np.pi * rand_b
    xx_t = np.broadcast_to(xx.reshape(1, n, xdim), (n_funcs, n, xdim))
    Z = np.sqrt(2.0 * sigma / n_features) * np.cos(np.matmul(W, np.transpose(xx_t, (0, 2, 1))) + b)
    noise = np.random.randn(n_funcs, n_features, 1) if draws is None else draws[2]
    yy_t = np.broadcast_to(yy.reshape(1, n, 1), (n_funcs, n, 1))
    Zt = np.transpose(Z, (0, 2, 1))
    if n < n_features:
        Sigma = np.matmul(Zt, Z) + sigma0 * np.eye(n)[None]
        mu = np.matmul(np.matmul(Z, np.linalg.inv(Sigma)), yy_t)
        e, v = np.linalg.eig(Sigma)
        e = e.reshape(n_funcs, n, 1)
        r = 1.0 / (np.sqrt(e) * (np.sqrt(e) + np.sqrt(sigma0)))
        theta = noise - np.matmul(Z, np.matmul(v, r * np.matmul(np.transpose(v, (0, 2, 1)),
                                                               np.matmul(Zt, noise)))) + mu
    else:
        Sigma = np.linalg.inv(np.matmul(Z, Zt) / sigma0 + np.eye(n_features)[None])
        mu = np.matmul(np.matmul(Sigma, Z), yy_t) / sigma0
        theta = mu + np.matmul(np.linalg.cholesky(Sigma), noise)
    return theta, W, b


def make_function_sample_np(x, n_features, sigma, theta, W, b):
    """optfunc.py:389-402: values of the function sample(s) at x -> (n_funcs, nx) or (nx,)."""
    theta = np.asarray(theta)
    if theta.ndim == 2:
        theta, W, b = theta[None], np.asarray(W)[None], np.asarray(b)[None]
    f = ops.rff_eval(x, W, b, theta, float(sigma))
    f = f.cpu().numpy() if isinstance(x, np.ndarray) else f
    return f[0] if f.shape[0] == 1 else f
