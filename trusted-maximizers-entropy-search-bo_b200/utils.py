"""Mirror of the reference's utils.py for the TES hot path (same function names and argument
meaning), backed by the CUDA kernels of libtes_b200.

Device work (SE-ARD kernels, Cholesky inverses, posteriors) goes through the C ABI.  The
trusted-maximizer bookkeeping of utils.py:1554-1833 is host-side numpy in the reference as well
(it runs between two sess.run calls); it is restated here with the random draw made injectable.
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
    """utils.py:657-674."""
    return _ret(ops.batched_chol2inv(mat), mat)


def multichol2inv(mat, n_mat=None, dtype=None):
    """utils.py:677-700."""
    return _ret(ops.batched_chol2inv(mat), mat)


def precomputeInvKs(xdim, nhyp, ls, sigmas, sigma0s, Xsamples, dtype=None):
    """utils.py:1838-1856 -> (nhyp, nobs, nobs)."""
    ls_, sg, s0 = np.asarray(_np(ls)).reshape(nhyp, xdim), _np(sigmas).reshape(-1), _np(sigma0s).reshape(-1)
    mats = torch.stack([dev(computeNKmm(dev(Xsamples), ls_[i], sg[i], s0[i])) for i in range(nhyp)])
    return _ret(ops.batched_chol2inv(mats), Xsamples)


def eval_invKmaxsams(xdim, nhyp, nmax, ls, sigmas, sigma0s, Xsamples, maxima, dtype=None):
    """utils.py:1859-1883: one (n+1)x(n+1) inverse per maximizer, batched on the device."""
    ls_, sg, s0 = np.asarray(_np(ls)).reshape(nhyp, xdim), _np(sigmas).reshape(-1), _np(sigma0s).reshape(-1)
    Xs = dev(Xsamples)
    mx = dev(maxima).reshape(nhyp, nmax, xdim)
    out = []
    for i in range(nhyp):
        stack = torch.stack([dev(computeNKmm(torch.cat([mx[i, j:j + 1], Xs], 0), ls_[i], sg[i], s0[i]))
                             for j in range(nmax)])
        out.append(ops.batched_chol2inv(stack))
    return _ret(torch.stack(out), Xsamples)


def compute_mean_var_f(x, Xsamples, Ysamples, l, sigma, sigma0, NKInv=None, fullcov=False, dtype=None):
    """utils.py:786-815."""
    xt, Xs = dev(x), dev(Xsamples)
    if NKInv is None:
        NKInv = ops.batched_chol2inv(dev(computeNKmm(Xs, l, sigma, sigma0)))
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


def evaluate_norm_entropy(s, dtype=None):
    """utils.py:653-654."""
    if torch.is_tensor(s):
        return 0.5 * np.log(2 * np.pi) + torch.log(s) + 0.5
    return 0.5 * np.log(2 * np.pi) + np.log(s) + 0.5


def _np(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


# ---- trusted-maximizer statistics (host side in the reference too) ---------------------------------
def get_duplicate_mask_np(xs, resolution=1e-5, return_leader=False):
    """utils.py:1554-1573: walking the rows in order, every not-yet-flagged row flags all LATER rows within
    `resolution` (Euclidean) as duplicates.  Same semantics, one pairwise-distance matrix instead of the
    O(n^2) Python loop; `leader[j]` is the row that flagged j (j itself for kept rows)."""
    xs = _np(xs)
    n = xs.shape[0]
    mask = np.zeros(n)
    leader = np.arange(n)
    if n:
        sq = np.sum(xs * xs, axis=1)
        # the reference computes sqrt(sum((xs[j]-xs[i])**2)); evaluate exactly that for the rows that matter
        close = (sq[:, None] + sq[None, :] - 2.0 * xs.dot(xs.T)) <= (resolution * 1.5) ** 2 + 1e-12
        for i in range(n):
            if mask[i] == 1.0:
                continue
            cand = np.nonzero(close[i, i + 1:])[0] + i + 1
            if cand.size:
                d = np.sqrt(np.sum((xs[cand] - xs[i]) ** 2, axis=1))
                hit = cand[(d <= resolution) & (mask[cand] == 0.0)]
                mask[hit] = 1.0
                leader[hit] = i
    return (mask, leader) if return_leader else mask


def remove_duplicates_np(xs, resolution=1e-5):
    """utils.py:1576-1581."""
    xs = _np(xs)
    return np.delete(xs, np.where(get_duplicate_mask_np(xs, resolution) == 1.0)[0], axis=0)


def mvn_transform_matrix(cov):
    """utils.py:1668-1670: A = V diag(sqrt(clip(e, 0))) from a GENERAL eig of the symmetric cov.  Column
    order and signs are LAPACK dgeev's and can change under 1e-13 perturbations of cov (near-degenerate
    eigenvalues), so callers that need identical samples for identical z pass this factor explicitly."""
    e, v = np.linalg.eig(_np(cov))
    return v.dot(np.diag(np.sqrt(np.clip(e, 0.0, np.inf))))


def _joint_sample_buckets(mean, cov, nsample, n_min_sample, remove_correlated_dims, z, transform):
    """Core of utils.py:1654-1729: joint samples, correlated-dimension suppression, arg-max bucketing.
    Returns (samples (nsample,n), unique maximizer ids, sizes, rows per bucket in sample order)."""
    mean, cov = _np(mean), _np(cov)
    n = cov.shape[0]
    mean = mean.reshape(n)
    A = mvn_transform_matrix(cov) if transform is None else _np(transform)
    if z is None:
        z = np.random.normal(size=(nsample, n, 1))
    # utils.py:1675 broadcasts mean.reshape(1,n) against the (nsample,n,1) product and then averages over the
    # LAST axis, so what is added to every coordinate is the AVERAGE of the prior means, not mean[i]
    # (reference quirk Q11, reproduced on purpose).  Evaluated here as ONE GEMM: z A^T + mean(mean).
    samples = _np(z).reshape(-1, n).dot(A.T) + np.mean(mean)
    corr_idxs = []
    if remove_correlated_dims:
        with np.errstate(invalid="ignore", divide="ignore"):
            rows, _ = np.where(np.tril(np.corrcoef(samples.T), k=-1) > 1 - 1e-6)
        corr_idxs = rows
    masked = samples
    if len(corr_idxs):
        masked = samples.copy()
        masked[:, corr_idxs] = -np.inf
    maxidxs = np.argmax(masked, axis=1)
    counts = np.bincount(maxidxs, minlength=n)
    unique = np.where(counts >= max(n_min_sample, 1))[0]
    sizes = counts[unique]
    order = np.argsort(maxidxs, kind="stable")
    starts = np.concatenate([[0], np.cumsum(counts)])
    rows_per = [order[starts[mi]:starts[mi + 1]] for mi in unique]
    return samples, mean, unique, sizes, rows_per


def sample_multivariate_normal_maxidx_np(mean, cov, nsample, n_min_sample=1, get_sample=True,
                                         remove_correlated_dims=True, z=None, transform=None):
    """utils.py:1654-1729.  `z` (nsample, n, 1) replaces the np.random.normal draw of :1674; when
    omitted the global numpy RNG is used exactly as the reference does.  `transform` overrides the
    colouring factor A (default: mvn_transform_matrix(cov), as the reference computes it)."""
    samples, mean, unique, sizes, rows_per = _joint_sample_buckets(mean, cov, nsample, n_min_sample,
                                                                   remove_correlated_dims, z, transform)
    if not get_sample:
        return unique, sizes
    n = mean.shape[0]
    K = int(sizes.max())
    nmax = len(unique)
    ret = np.tile(mean.reshape(1, n, 1), (nmax, 1, K))  # padded slots hold the prior mean (:1722)
    masks = np.zeros([nmax, K], dtype=int)
    for i in range(nmax):
        ret[i, :, :sizes[i]] = samples[rows_per[i]].T
        masks[i, :sizes[i]] = 1
    return nmax, unique, K, ret, masks, sizes


def get_testidxs_stats_host(nhyp, test_xs_np, test_means, test_covs, mode="sample", nsample=1000, n_min_sample=2,
                            z=None, transforms=None):
    """Host-numpy restatement of utils.py:1733-1833 (where the reference runs it).  Kept for mode='ep' (whose
    EP sweeps are host code here as in the reference) and as the numpy-level mirror; the default entry point
    `get_testidxs_stats` below runs modes 'sample' and 'empirical' on the device.
    mode in {'sample','empirical','ep'}; `z` (nhyp, nsample, ntest, 1) optional.
    mode='ep' runs the EP the reference intends at :1820-1825 (its shipped code raises NameError).
    The padded (nhyp, ntest, ntest, K) sample tensor is only materialised for mode='sample'; the empirical
    moments are taken straight from the buckets (zero-weight padding contributes nothing)."""
    from . import empirical_approximation, ep
    test_xs_np, test_means, test_covs = _np(test_xs_np), _np(test_means), _np(test_covs)
    if mode not in ("sample", "empirical", "ep"):
        raise Exception("Unknown mode in get_testidxs_stats!")
    ntest = test_xs_np.shape[0]
    alive = np.ones(ntest, dtype=bool)
    max_probs = np.ones([nhyp, ntest]) / ntest
    per_hyp = []
    for i in range(nhyp):
        samples, mean_i, uniq, sizes, rows_per = _joint_sample_buckets(
            test_means[i], test_covs[i], nsample, n_min_sample, True, None if z is None else z[i],
            None if transforms is None else transforms[i])
        won = np.zeros(ntest, dtype=bool)
        won[uniq] = True
        alive &= won
        max_probs[i, uniq] = sizes / np.sum(sizes)
        per_hyp.append((samples, mean_i, uniq, sizes, rows_per))
    keep = np.where(alive)[0]
    pos = -np.ones(ntest, dtype=int)
    pos[keep] = np.arange(len(keep))
    if len(keep) != ntest:
        test_xs_np = test_xs_np[keep]
        test_means = test_means[:, keep]
        test_covs = test_covs[:, keep][:, :, keep]
        max_probs = max_probs[:, keep]
    max_probs = max_probs / np.sum(max_probs, axis=1, keepdims=True)
    nk = len(keep)
    if mode == "sample":
        K_all = max(int(sz.max()) for (_, _, _, sz, _) in per_hyp)
        post_samples = np.zeros([nhyp, nk, nk, K_all])
        post_masks = np.zeros([nhyp, nk, K_all], dtype=bool)
        for i, (samples, mean_i, uniq, sizes, rows_per) in enumerate(per_hyp):
            K = int(sizes.max())
            for u, sz, rows in zip(uniq, sizes, rows_per):
                if pos[u] < 0:
                    continue
                post_samples[i, pos[u], :, :sz] = samples[rows][:, keep].T
                post_samples[i, pos[u], :, sz:K] = mean_i[keep].reshape(nk, 1)  # prior-mean padding (:1722)
                post_masks[i, pos[u], :sz] = True
        return test_xs_np, max_probs, test_means, test_covs, post_samples, post_masks
    pm = np.zeros([nhyp, nk, nk])
    pc = np.zeros([nhyp, nk, nk, nk])
    for i, (samples, mean_i, uniq, sizes, rows_per) in enumerate(per_hyp):
        if mode == "empirical":
            sel = [rows for u, rows in zip(uniq, rows_per) if pos[u] >= 0]
            pm[i], pc[i] = empirical_approximation.bucket_empirical_stats(samples[:, keep], sel)
        else:
            for j in range(nk):
                m, c = ep.approximate_EP_host(j, test_means[i].reshape(-1, 1), test_covs[i], resolution=1e-9,
                                            max_niter=100)
                pm[i, j] = m.squeeze()
                pc[i, j] = c
    return test_xs_np, max_probs, test_means, test_covs, pm, pc


def _device_buckets(mean, cov, nsample, z, transform, seed, hyp_index):
    """utils.py:1668-1705 on the device: colouring factor (Jacobi eigen-solver unless `transform` is given),
    joint samples z A^T + mean(mean) (Q11), correlated-dimension mask, row arg-max, bucket counts."""
    T = cov.shape[0]
    A = ops.color_factor(cov) if transform is None else dev(transform, cov.device)
    if z is not None:
        zt = dev(z, cov.device).reshape(-1, T)
    elif seed is not None:
        zt = ops.normal_fill(seed, hyp_index, ops.STREAM_JOINT_Z, nsample * T, device=cov.device).reshape(nsample, T)
    else:  # the reference's own draw: global numpy RNG (utils.py:1674)
        zt = dev(np.random.normal(size=(nsample, T, 1)), cov.device).reshape(nsample, T)
    samples = ops.gemm(zt, A, transb=True)
    _, maxidx, counts = ops.joint_maxidx(samples, mean.reshape(-1), True)
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
