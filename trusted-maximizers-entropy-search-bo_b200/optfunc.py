"""Mirror of the reference's optfunc.py (numpy half): RFF posterior weight draw and function-sample
evaluation, both on the device (the draw is host numpy in the reference, bo_batch.py:641-652)."""
import numpy as np
import torch

from . import ops
from .device import dev


def draw_random_init_weights_features(xdim, n_funcs, n_features, xx, yy, l, sigma, sigma0, draws=None, seed=None,
                                      hyp_index=0):
    """optfunc.py:303-385 on the device (tes_rff_draw_f64).  The three random draws are, in this order of
    precedence: `draws` = (randn_W (S,F,d), rand_b (S,F,1), randn_noise (S,F,1)) host- or device-supplied;
    `seed` -> counter-based draws generated on the device (include/tes_spec.h streams TES_STREAM_RFF_W/B/E);
    neither -> np.random in the reference's order (:329, :334, :344) copied to the device.
    Returns CUDA tensors theta (S,F,1), W (S,F,d), b (S,F,1)."""
    S, F, d = int(n_funcs), int(n_features), int(xdim)
    if draws is not None:
        randn_W, rand_b, noise = draws
    elif seed is not None:
        randn_W = ops.normal_fill(seed, hyp_index, ops.STREAM_RFF_W, S * F * d).reshape(S, F, d)
        rand_b = ops.uniform_fill(seed, hyp_index, ops.STREAM_RFF_B, S * F).reshape(S, F)
        noise = ops.normal_fill(seed, hyp_index, ops.STREAM_RFF_E, S * F).reshape(S, F)
    else:
        randn_W = np.random.randn(S, F, d)
        rand_b = np.random.rand(S, F, 1)
        noise = np.random.randn(S, F, 1)
    theta, W, b = ops.rff_draw(dev(xx).reshape(-1, d), yy, l, sigma, sigma0, randn_W, rand_b, noise)
    return theta.reshape(S, F, 1), W, b.reshape(S, F, 1)


def draw_random_init_weights_features_np(xdim, n_funcs, n_features, xx, yy, l, sigma, sigma0, draws=None, seed=None):
    """optfunc.py:303-385, numpy in -> numpy out (the reference's signature), computed on the device.
    `draws` = (randn_W (S,F,d), rand_b (S,F,1), randn_noise (S,F,1)) makes the three global-RNG draws
    (:329, :334, :344) explicit; omitted -> np.random in the same order."""
    theta, W, b = draw_random_init_weights_features(xdim, n_funcs, n_features, xx, yy, l, sigma, sigma0, draws, seed)
    return theta.cpu().numpy(), W.cpu().numpy(), b.cpu().numpy()


def make_function_sample_np(x, n_features, sigma, theta, W, b):
    """optfunc.py:389-402: values of the function sample(s) at x -> (n_funcs, nx) or (nx,)."""
    theta = np.asarray(theta)
    if theta.ndim == 2:
        theta, W, b = theta[None], np.asarray(W)[None], np.asarray(b)[None]
    f = ops.rff_eval(x, W, b, theta, float(sigma))
    f = f.cpu().numpy() if isinstance(x, np.ndarray) else f
    return f[0] if f.shape[0] == 1 else f
