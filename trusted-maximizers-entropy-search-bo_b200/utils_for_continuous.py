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
