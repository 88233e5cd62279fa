"""Mirror of the reference's utils.py for the TES hot path (same function names and argument
meaning), backed by the CUDA kernels of libtes_b200.

All numerical work (SE-ARD kernels, Cholesky inverses, posteriors, and the trusted-maximizer statistics of
utils.py:1554-1833 that the reference runs as host numpy between two sess.run calls) goes through the C ABI;
only bucket COUNTS visit the host (they size the outputs).  The random draws are injectable.  There is no host
implementation in this package (the numpy restatement the tests check against is test infrastructure, outside this package).
Arrays may be numpy or torch; numpy in -> numpy out.
"""
import numpy as np
import torch

from . import ops
from .device import dev

clip_min = 1e-100  # utils.py:11


def _ret(t, like):
    return t.cpu().numpy() if isinstance(like, np.ndarray) else t


# ---- kernels / posteriors ------------------------------------------------------------------------
def computeKnm(X, Xbar, l, sigma, dtype=None):
    """utils.py:703-719."""
    return _ret(ops.se_ard_cross(X, Xbar, np.asarray(l).reshape(-1) if not torch.is_tensor(l) else l.reshape(-1),
                                 float(sigma)), X)


computeKnm_np = computeKnm  # utils.py:836-859


def computeKmm(X, l, sigma, nd=2, dtype=None):
    """utils.py:740-762 (nd == 2) -- Gram matrix."""
    Xt = dev(X)
    if Xt.dim() != 2:
        return _ret(torch.stack([ops.se_ard_cross(x, x, l, float(sigma)) for x in Xt.reshape(-1, *Xt.shape[-2:])])
                    .reshape(*Xt.shape[:-1], Xt.shape[-2]), X)
    return _ret(ops.se_ard_cross(Xt, Xt, l, float(sigma)), X)


computeKmm_np = computeKmm  # utils.py:818-833


def computeNKmm(X, l, sigma, sigma0, dtype=None):
    """utils.py:765-783 (no jitter)."""
    Xt = dev(X)
    K = ops.se_ard_cross(Xt, Xt, l, float(sigma))
    K.diagonal().add_(float(sigma0))
    return _ret(K, X)


def chol2inv(mat, dtype=None):
    """utils.py:657-674.  Raises TesError where tf.linalg.cholesky would fail."""
    return _ret(ops.chol2inv_checked(mat, "chol2inv"), mat)


def multichol2inv(mat, n_mat=None, dtype=None):
    """utils.py:677-700.  Raises TesError where tf.linalg.cholesky would fail."""
    return _ret(ops.chol2inv_checked(mat, "multichol2inv"), mat)


def precomputeInvKs(xdim, nhyp, ls, sigmas, sigma0s, Xsamples, dtype=None):
    """utils.py:1838-1856 -> (nhyp, nobs, nobs)."""
    ls_, sg, s0 = np.asarray(_np(ls)).reshape(nhyp, xdim), _np(sigmas).reshape(-1), _np(sigma0s).reshape(-1)
    mats = torch.stack([dev(computeNKmm(dev(Xsamples), ls_[i], sg[i], s0[i])) for i in range(nhyp)])
    return _ret(ops.chol2inv_checked(mats, "precomputeInvKs"), Xsamples)


def eval_invKmaxsams(xdim, nhyp, nmax, ls, sigmas, sigma0s, Xsamples, maxima, dtype=None):
    """utils.py:1859-1883: one (n+1)x(n+1) inverse per maximizer, batched on the device."""
    ls_, sg, s0 = np.asarray(_np(ls)).reshape(nhyp, xdim), _np(sigmas).reshape(-1), _np(sigma0s).reshape(-1)
    Xs = dev(Xsamples)
    mx = dev(maxima).reshape(nhyp, nmax, xdim)
    out = []
    for i in range(nhyp):
        stack = torch.stack([dev(computeNKmm(torch.cat([mx[i, j:j + 1], Xs], 0), ls_[i], sg[i], s0[i]))
                             for j in range(nmax)])
        out.append(ops.chol2inv_checked(stack, "eval_invKmaxsams"))
    return _ret(torch.stack(out), Xsamples)


def compute_mean_var_f(x, Xsamples, Ysamples, l, sigma, sigma0, NKInv=None, fullcov=False, dtype=None):
    """utils.py:786-815."""
    xt, Xs = dev(x), dev(Xsamples)
    if NKInv is None:
        NKInv = ops.chol2inv_checked(dev(computeNKmm(Xs, l, sigma, sigma0)), "compute_mean_var_f")
    NKInv = dev(NKInv)
    if not fullcov:
        mean, var = ops.gp_mean_var(xt, Xs, Ysamples, l, float(sigma), NKInv)
        return _ret(mean, x), _ret(var, x)
    kstar = ops.se_ard_cross(xt, Xs, l, float(sigma))
    mean = ops.gemm(kstar, ops.gemm(NKInv, dev(Ysamples).reshape(-1, 1))).reshape(-1)
    var = ops.se_ard_cross(xt, xt, l, float(sigma)) - ops.gemm(ops.gemm(kstar, NKInv), kstar, transb=True)
    var.diagonal().clamp_(min=clip_min)
    return _ret(mean, x), _ret(var, x)


def compute_mean_f(x, xdim, n_hyp, Xsamples, Ysamples, ls, sigmas, sigma0s, NKInvs, dtype=None):
    """utils.py:982-1009."""
    ls_, sg = np.asarray(_np(ls)).reshape(n_hyp, xdim), _np(sigmas).reshape(-1)
    acc = 0.0
    for i in range(n_hyp):
        mean, _ = ops.gp_mean_var(x, Xsamples, Ysamples, ls_[i], float(sg[i]), dev(NKInvs)[i])
        acc = acc + mean / n_hyp
    return _ret(acc, x)


NORM_ENTROPY_CONST = 0.918938577175140380859375  # float32(0.5) * logf(float32(2*pi))


def evaluate_norm_entropy(s, dtype=None):
    """utils.py:653-654.  The reference's constant is tf.cast(0.5 * tf.log(2*np.pi), dtype): a bare Python
    float becomes a float32 tensor in TF 1.14, so it is float32(0.5*logf(2*pi)) widened (quirk Q12)."""
    c = NORM_ENTROPY_CONST
    if torch.is_tensor(s):
        return c + torch.log(s) + 0.5
    return c + np.log(s) + 0.5


def _np(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


# ---- trusted-maximizer statistics (host numpy in the reference; device kernels here) ----------------
def get_duplicate_mask_np(xs, resolution=1e-5, return_leader=False):
    """utils.py:1554-1573 on the device (tes_dedup_f64): walking the rows in order, every not-yet-flagged row
    flags all LATER rows within `resolution` (Euclidean) as duplicates.  Returns the reference's float mask
    (1.0 = duplicate); `leader[j]` is the row that flagged j (j itself for kept rows)."""
    mask, leader = ops.dedup(dev(xs), resolution)
    mask = mask.to(torch.float64)
    if not torch.is_tensor(xs):
        mask, leader = mask.cpu().numpy(), leader.cpu().numpy()
    return (mask, leader) if return_leader else mask


def remove_duplicates_np(xs, resolution=1e-5):
    """utils.py:1576-1581 on the device: the rows of xs that are not flagged as duplicates, original order."""
    xs_d = dev(xs)
    mask, _ = ops.dedup(xs_d, resolution)
    return _ret(xs_d[mask == 0], xs)


def sample_multivariate_normal_maxidx_np(mean, cov, nsample, n_min_sample=1, get_sample=True,
                                         remove_correlated_dims=True, z=None, transform=None, seed=None):
    """utils.py:1654-1729 on the device.  `z` (nsample, n, 1) replaces the np.random.normal draw of :1674
    (`seed`: counter-based draws; neither: the global numpy RNG exactly as the reference).  `transform`
    overrides the colouring factor A of :1668-1670 (default: the device eigen-solver's factor -- same law as
    LAPACK's, different realisation).  Returns (unique, sizes) or
    (nmax, unique, K, samples (nmax,n,K), masks (nmax,K), sizes) as the reference."""
    like = mean
    mean_d, cov_d = dev(mean).reshape(-1), dev(cov)
    n = cov_d.shape[0]
    samples, maxidx, counts = _device_buckets(mean_d, cov_d, nsample, z, transform, seed, 0,
                                              remove_correlated=remove_correlated_dims)
    counts_h = counts.cpu().numpy().astype(np.int64)
    unique = np.where(counts_h >= max(n_min_sample, 1))[0]
    sizes = counts_h[unique]
    if not get_sample:
        return unique, sizes
    K = int(sizes.max())
    order = torch.sort(maxidx, stable=True).indices
    starts_all = np.concatenate([[0], np.cumsum(counts_h)])
    pick = np.concatenate([np.arange(starts_all[u], starts_all[u + 1]) for u in unique])
    order_k = order[torch.from_numpy(pick).to(cov_d.device)].contiguous()
    starts_k = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)).to(cov_d.device)
    cols = torch.arange(n, dtype=torch.int64, device=cov_d.device)
    ret, masks = ops.bucket_samples(samples, order_k, starts_k, cols, K, mean_d)   # prior-mean padding (:1722)
    masks = masks.to(torch.int64)
    if not torch.is_tensor(like):
        ret, masks = ret.cpu().numpy(), masks.cpu().numpy()
    return len(unique), unique, K, ret, masks, sizes


def _device_buckets(mean, cov, nsample, z, transform, seed, hyp_index, remove_correlated=True):
    """utils.py:1668-1705 on the device: colouring factor (`transform`, else Cholesky, else the Jacobi eigen-solver),
    joint samples z A^T + mean(mean) (Q11), correlated-dimension mask, row arg-max, bucket counts."""
    T = cov.shape[0]
    if transform is not None:
        A = dev(transform, cov.device)
    else:
        # any A with A A^T = cov gives the reference's law.  The Cholesky factor (blocked potrf, ~0.3 ms at T = 128)
        # is used when cov is numerically positive definite; otherwise the reference's own construction
        # V sqrt(clip(e, 0)) from the device eigen-solver (12 ms at T = 128), which also covers semi-definite cov.
        A, info = ops.batched_potrf(cov, return_info=True)
        if int(info.reshape(-1)[0]) != 0 or not bool(torch.isfinite(A).all()):
            A = ops.color_factor(cov)
    if z is not None:
        zt = dev(z, cov.device).reshape(-1, T)
    elif seed is not None:
        zt = ops.normal_fill(seed, hyp_index, ops.STREAM_JOINT_Z, nsample * T, device=cov.device).reshape(nsample, T)
    else:  # the reference's own draw: global numpy RNG (utils.py:1674)
        zt = dev(np.random.normal(size=(nsample, T, 1)), cov.device).reshape(nsample, T)
    samples = ops.gemm(zt, A, transb=True)
    _, maxidx, counts = ops.joint_maxidx(samples, mean.reshape(-1), remove_correlated)
    return samples, maxidx, counts


def get_testidxs_stats(nhyp, test_xs_np, test_means, test_covs, mode="sample", nsample=1000, n_min_sample=2,
                       z=None, transforms=None, seed=None):
    """utils.py:1733-1833 with the sampling, bucketing and moment estimation on the device.
    mode in {'sample','empirical','ep'}; `z` (nhyp, nsample, ntest, 1) replaces the np.random.normal draw of
    utils.py:1674 (or `seed`: counter-based draws of include/tes_spec.h, stream TES_STREAM_JOINT_Z);
    `transforms[i]` overrides the colouring factor of hyper-parameter sample i.  numpy in -> numpy out, CUDA
    tensors in -> CUDA tensors out.  Only the T bucket counts cross to the host (they size the outputs)."""
    if mode not in ("sample", "empirical", "ep"):
        raise Exception("Unknown mode in get_testidxs_stats!")
    like = test_means
    xs_d, means_d, covs_d = dev(test_xs_np), dev(test_means), dev(test_covs)
    dv = means_d.device
    ntest = xs_d.shape[0]
    means_d, covs_d = means_d.reshape(nhyp, ntest), covs_d.reshape(nhyp, ntest, ntest)
    means_full = means_d
    alive = np.ones(ntest, dtype=bool)
    max_probs = np.ones([nhyp, ntest]) / ntest
    per_hyp = []
    for i in range(nhyp):
        samples, maxidx, counts = _device_buckets(means_d[i], covs_d[i], nsample, None if z is None else z[i],
                                                  None if transforms is None else transforms[i], seed, i)
        counts_h = counts.cpu().numpy().astype(np.int64)
        uniq = np.where(counts_h >= max(n_min_sample, 1))[0]
        sizes = counts_h[uniq]
        won = np.zeros(ntest, dtype=bool)
        won[uniq] = True
        alive &= won
        max_probs[i, uniq] = sizes / np.sum(sizes)
        per_hyp.append((samples, maxidx, counts_h, uniq, sizes))
    keep = np.where(alive)[0]
    nk = len(keep)
    keep_d = torch.from_numpy(keep).to(dv)
    if nk != ntest:
        xs_d = xs_d[keep_d]
        means_d = means_d[:, keep_d]
        covs_d = covs_d[:, keep_d][:, :, keep_d]
        max_probs = max_probs[:, keep]
    max_probs = max_probs / np.sum(max_probs, axis=1, keepdims=True)
    head = (_ret(xs_d, like), _ret(dev(max_probs, dv), like), _ret(means_d, like), _ret(covs_d, like))
    if mode == "ep":
        if nk < 2:  # a single surviving test point has no half-space sites: the prior is the answer
            return head + (_ret(means_d.reshape(nhyp, nk, nk), like), _ret(covs_d.reshape(nhyp, nk, nk, nk), like))
        pm, pc = ops.ep_moments(means_d, covs_d, resolution=1e-9, max_niter=100)   # utils.py:1820-1825
        return head + (_ret(pm, like), _ret(pc, like))
    K_all = max([int(sz.max()) if len(sz) else 1 for (_, _, _, _, sz) in per_hyp])
    outs0, outs1 = [], []
    for i, (samples, maxidx, counts_h, uniq, sizes) in enumerate(per_hyp):
        order = torch.sort(maxidx, stable=True).indices          # sample rows grouped by bucket, draw order inside
        starts_all = np.concatenate([[0], np.cumsum(counts_h)])
        kept_b = [u for u in uniq if alive[u]]                    # buckets of the kept test points, ascending
        pick = np.concatenate([np.arange(starts_all[u], starts_all[u + 1]) for u in kept_b]) if kept_b else \
            np.zeros(0, dtype=np.int64)
        order_k = order[torch.from_numpy(pick).to(dv)].contiguous()
        starts_k = torch.from_numpy(np.concatenate([[0], np.cumsum(counts_h[kept_b])]).astype(np.int64)).to(dv)
        if mode == "sample":
            ps, pmask = ops.bucket_samples(samples, order_k, starts_k, keep_d, K_all, means_full[i])
            # the reference pads every hyper-parameter sample only up to its own K with the prior mean and leaves
            # zeros beyond (utils.py:1797-1809)
            Ki = int(sizes.max())
            if Ki < K_all:
                ps[:, :, Ki:] = 0.0
            outs0.append(ps)
            outs1.append(pmask)
        else:
            bm, bc = ops.bucket_moments(samples, order_k, starts_k, keep_d)
            outs0.append(bm)
            outs1.append(bc)
    return head + (_ret(torch.stack(outs0), like), _ret(torch.stack(outs1), like))
