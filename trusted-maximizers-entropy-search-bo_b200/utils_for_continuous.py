"""Mirror of utils_for_continuous.sample_xmaxs_fmaxs (utils_for_continuous.py:138-252): the hot loop
:172-187 (all function samples on all initializers + top-k) runs as one CUDA kernel, and the Adam refinement of
:201-232 (which the reference drives with `ntrain` sess.run(train) calls, bo_batch.py:676-690) as one more launch when
`ntrain` > 0."""
import numpy as np
import torch

from . import ops
from .device import dev


def sample_xmaxs_fmaxs(xdim, n_hyp, n_funcs, n_features, ls, sigmas, sigma0s, thetas_all, Ws_all, bs_all,
                       initializers, top_init_k, xmin=None, xmax=None, dtype=None, parallel_iterations=None,
                       get_xs=False, name=None, idx_offset=0, ntrain=0):
    """Returns (top_k_inits (nhyp,n_funcs,k,xdim), max_xs (nhyp,n_funcs,xdim), max_fs (nhyp,n_funcs),
    top_idx (nhyp,n_funcs,k)).  ntrain == 0: the arg-max is the first of the top-k (tf.math.top_k is sorted).
    ntrain > 0: the top-k points are refined by `ntrain` Adam steps inside [xmin, xmax] (:201-220) and
    max_xs / max_fs are the best refined point of every sample (:222-232); with get_xs the refined
    (nhyp,n_funcs,k,xdim) points are appended to the result."""
    init = dev(initializers)
    sg = np.asarray(sigmas if not torch.is_tensor(sigmas) else sigmas.cpu()).reshape(-1)
    inits, mx, mf, ti, xs_all = [], [], [], [], []
    for i in range(n_hyp):
        idx, val = ops.rff_topk(init, dev(Ws_all)[i], dev(bs_all)[i], dev(thetas_all)[i], float(sg[i]),
                                k=top_init_k, idx_offset=idx_offset)
        local = (idx - idx_offset).clamp_(min=0)
        pts = init[local.reshape(-1)].reshape(idx.shape[0], top_init_k, xdim)
        inits.append(pts)
        ti.append(idx)
        if ntrain > 0:
            xr, fr = ops.rff_adam_refine(pts, dev(Ws_all)[i], dev(bs_all)[i], dev(thetas_all)[i], float(sg[i]),
                                         0.0 if xmin is None else xmin, 1.0 if xmax is None else xmax, ntrain)
            best = torch.argmax(fr, dim=1)                      # first maximum, as tf.argmax
            rows = torch.arange(fr.shape[0], device=fr.device)
            mx.append(xr[rows, best])
            mf.append(fr[rows, best])
            xs_all.append(xr)
        else:
            mx.append(pts[:, 0])
            mf.append(val[:, 0])
            xs_all.append(pts)
    out = (torch.stack(inits), torch.stack(mx), torch.stack(mf), torch.stack(ti))
    if get_xs:
        out = out + (torch.stack(xs_all),)
    if isinstance(initializers, np.ndarray):
        return tuple(o.cpu().numpy() for o in out)
    return out


def adam_ascent(func, x0, xmin, xmax, ntrain, fd_step=1e-5, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """`ntrain` TF-1 Adam steps (tf.train.AdamOptimizer defaults; lr_t = lr sqrt(1-b2^t)/(1-b1^t),
    x += lr_t m / (sqrt(v) + eps)) maximising sum_i func(x)[i] over the rows of x (k, D), each update followed by the
    variable's clip constraint (utils_for_continuous.py:62-77).  The reference differentiates its TF graph; here the
    gradient is a central difference with step `fd_step` * (xmax - xmin): func is called ONCE per step on the
    (k (2 D + 1), D) stack of perturbed points -- one batched criterion launch -- and must use common random numbers
    across the rows of a call (step index passed as func(x, step)).  Returns (x (k, D), f(x) (k,))."""
    x = np.array(x0, dtype=np.float64, copy=True)
    k, D = x.shape
    lo = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (D,))
    hi = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (D,))
    h = fd_step * np.where(np.isfinite(hi - lo), hi - lo, 1.0)
    m, v = np.zeros_like(x), np.zeros_like(x)
    eye = np.eye(D)
    for t in range(1, ntrain + 1):
        pts = np.concatenate([x[:, None, :] + h * eye[None], x[:, None, :] - h * eye[None]], axis=1)   # (k, 2D, D)
        f = np.asarray(func(pts.reshape(-1, D), t), dtype=np.float64).reshape(k, 2 * D)
        g = -(f[:, :D] - f[:, D:]) / (2.0 * h)               # gradient of the minimised -sum f
        m = beta1 * m + (1.0 - beta1) * g
        v = beta2 * v + (1.0 - beta2) * g * g
        lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
        x = np.clip(x - lr_t * m / (np.sqrt(v) + epsilon), lo, hi)
    return x, np.asarray(func(x, ntrain + 1), dtype=np.float64).reshape(k)


def optimize_continuous_function(xdim, func, candidate_xs, top_init_k, xmin=-np.inf, xmax=np.inf, ntrain=0,
                                 rand_candidate_xs=None, fd_step=1e-5, debug=False):
    """utils_for_continuous.optimize_continuous_function (:20-87) re-hosted: `top_init_k` uniform random candidates are
    appended to `candidate_xs` (:36-38; np.random.rand here, or `rand_candidate_xs` to inject them), every candidate is
    scored, the top_init_k best start `ntrain` Adam steps inside [xmin, xmax] (the sess.run(train) loop of the
    drivers, e.g. bo_batch.py:1048-1050), and the best refined point is returned with its value (first maximum).
    func(x (n, xdim), step) -> (n,) runs on the device; see adam_ascent for the gradient.
    Returns (maximizer (1, xdim), maximum) or, with debug, also (xs (k, xdim), fs (k,))."""
    cand = np.asarray(candidate_xs, dtype=np.float64).reshape(-1, xdim)
    if rand_candidate_xs is None:
        lo = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (xdim,))
        hi = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (xdim,))
        rand_candidate_xs = lo + np.random.rand(top_init_k, xdim) * (hi - lo)
    cand = np.concatenate([cand, np.asarray(rand_candidate_xs, dtype=np.float64).reshape(-1, xdim)], axis=0)
    fs0 = np.asarray(func(cand, 0), dtype=np.float64).reshape(-1)
    order = np.lexsort((np.arange(fs0.shape[0]), -fs0))[:top_init_k]      # tf.math.top_k: value desc, index asc
    xs, fs = adam_ascent(func, cand[order], xmin, xmax, ntrain, fd_step=fd_step)
    best = int(np.argmax(fs))
    out = (xs[best].reshape(1, xdim), float(fs[best]))
    return out + (xs, fs) if debug else out
