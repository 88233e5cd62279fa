"""Mirror of utils_for_continuous.sample_xmaxs_fmaxs (utils_for_continuous.py:138-252): the hot loop
:172-187 (all function samples on all initializers + top-k) runs as one CUDA kernel.  The Adam
refinement of :201-232 is a driver-level follow-up and is not part of this call."""
import numpy as np
import torch

from . import ops
from .device import dev


def sample_xmaxs_fmaxs(xdim, n_hyp, n_funcs, n_features, ls, sigmas, sigma0s, thetas_all, Ws_all, bs_all,
                       initializers, top_init_k, xmin=None, xmax=None, dtype=None, parallel_iterations=None,
                       get_xs=False, name=None, idx_offset=0):
    """Returns (top_k_inits (nhyp,n_funcs,k,xdim), max_xs (nhyp,n_funcs,xdim), max_fs (nhyp,n_funcs),
    top_idx (nhyp,n_funcs,k)) -- the arg-max is the first of the top-k (tf.math.top_k is sorted)."""
    init = dev(initializers)
    sg = np.asarray(sigmas if not torch.is_tensor(sigmas) else sigmas.cpu()).reshape(-1)
    inits, mx, mf, ti = [], [], [], []
    for i in range(n_hyp):
        idx, val = ops.rff_topk(init, dev(Ws_all)[i], dev(bs_all)[i], dev(thetas_all)[i], float(sg[i]),
                                k=top_init_k, idx_offset=idx_offset)
        local = (idx - idx_offset).clamp_(min=0)
        pts = init[local.reshape(-1)].reshape(idx.shape[0], top_init_k, xdim)
        inits.append(pts)
        mx.append(pts[:, 0])
        mf.append(val[:, 0])
        ti.append(idx)
    out = (torch.stack(inits), torch.stack(mx), torch.stack(mf), torch.stack(ti))
    if isinstance(initializers, np.ndarray):
        return tuple(o.cpu().numpy() for o in out)
    return out
