"""TEST INFRASTRUCTURE ONLY -- CPU fp64 numpy restatement of the reference's TES hot path.

This file is the *oracle*: a plain-numpy restatement of the algorithm the reference
(ZhaoxuanWu/Trusted-Maximizers-Entropy-Search-BO) implements with TensorFlow 1.14 graph ops
+ TFP 0.7 (criteria/*.py, utils.py, utils_for_continuous.py) and with numpy/scipy
(optfunc.py, ep.py, empirical_approximation.py, utils.py host half).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it;
the product package never does (it fails loudly when the CUDA library is missing).

Every function cites the reference file:line it follows.  Every random draw the reference
makes *inside* a TF graph or from the global numpy RNG is an explicit argument here
("Random draws consumed by the hot path").

Pinning:
  * computeKnm/computeKmm/compute_mean_f are pinned by the reference's only known-answer
    data, the Brooms-Barn objective maxima (functions.py:397-398, :363-364, :329-330);
    see tests/test_oracle_golden.py.
  * rff_fvals, draw_random_init_weights_features, sample_multivariate_normal_maxidx,
    get_testidxs_stats, get_empirical_stat_from_samples and approximate_EP are pinned against
    the outputs of the reference's own numpy functions, run in the build container through
    oracle/ref_loader.py by tests/golden/make_golden.py (fixtures in tests/golden/*.npz).
  * The five TF criteria (evaluate_mp, evaluate_mp_lite, evaluate_sample_mp, evaluate_batch_sample_mp,
    evaluate_batch_mp), get_pNK_test_obs and the TF halves of utils.py (precomputeInvKs, compute_mean_var_f) are
    pinned on outputs of the reference's OWN files, run unmodified on a numpy-backed TF-1.14/TFP-0.7 shim
    (oracle/tf_shim; tests/golden/make_golden_criteria.py -> tests/golden/criteria.npz, C1 / C2 / d=6 / nhyp=2 shapes
    with the draws its TFP sample() calls consumed): tests/test_oracle_criteria_golden.py.  Running them that way
    exposed quirk Q12 (the float32-rounded constant of utils.evaluate_norm_entropy).
  * adam_refine_maximizers is pinned on utils_for_continuous.sample_xmaxs_fmaxs run unmodified on a lazy-graph shim
    (oracle/tf_shim/graph.py: tf.Variable / assign / tf.train.AdamOptimizer per TF 1.14 with autograd gradients
    through the reference's graph; tests/golden/make_golden_adam.py -> tests/golden/adam.npz).
  What the shims restate is TensorFlow's published op semantics (third-party, tensorflow-gpu 1.14.0 /
  tensorflow-probability 0.7.0 per requirements.txt:5-6, not vendored), not the reference's algorithm.

TFP-0.7 semantics used (not vendored in /root/reference; requirements.txt:5-6):
  Normal.sample(n)        = loc + scale * z,  z ~ N(0,1) of shape (n,) + batch_shape
  Normal.log_prob(x)      = -0.5*((x-loc)/scale)**2 - (0.5*log(2*pi) + log(scale))
  MultivariateNormalFullCovariance(loc, cov) = TriL with L = cholesky(cov):
      sample = loc + L z ;  log_prob(x) = -0.5*||L^-1 (x-loc)||^2 - sum(log diag L) - q/2 log(2 pi)
  tf.reduce_logsumexp     = max-shifted, a non-finite max is replaced by 0 before the shift
"""
import math

import numpy as np
import scipy.stats as spst

CLIP_MIN = 1e-100  # utils.py:11
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
# utils.py:653-654 computes its constant as tf.cast(0.5 * tf.log(2*np.pi), dtype): TF 1.14 converts the bare
# Python float to a float32 tensor, so the value is float32(0.5) * logf(float32(2*pi)) widened to fp64
# (quirk Q12; found by running the reference's own evaluate_mp.py on oracle/tf_shim)
NORM_ENTROPY_CONST = float(np.float32(0.5) * np.log(np.float32(2.0 * np.pi)))
assert NORM_ENTROPY_CONST == 0.918938577175140380859375


# --------------------------------------------------------------------------------------
# GP kernels and posteriors  (utils.py:657-885, 1838-1856)
# --------------------------------------------------------------------------------------
def computeKnm(X, Xbar, l, sigma):
    """SE-ARD cross kernel, utils.py:703-719 (numpy twin :836-859).

    `l` is an INVERSE SQUARED length-scale: k = sigma*exp(-0.5*sum_d l_d (x-x')^2),
    evaluated as |x sqrt(l)|^2 + |x' sqrt(l)|^2 - 2 (x sqrt(l)).(x' sqrt(l)).
    """
    X = np.asarray(X, dtype=np.float64)
    Xbar = np.asarray(Xbar, dtype=np.float64)
    l = np.asarray(l, dtype=np.float64).reshape(1, -1)
    Xs = X * np.sqrt(l)
    Xbs = Xbar * np.sqrt(l)
    Q = np.sum(Xs * Xs, axis=1, keepdims=True)
    Qbar = np.sum(Xbs * Xbs, axis=1, keepdims=True).T
    dist = Qbar + Q - 2.0 * Xs.dot(Xbs.T)
    return sigma * np.exp(-0.5 * dist)


def computeKmm(X, l, sigma):
    """SE-ARD Gram matrix over the last two dims, utils.py:740-762 (numpy twin :818-833)."""
    X = np.asarray(X, dtype=np.float64)
    l = np.asarray(l, dtype=np.float64).reshape(1, -1)
    Xs = X * np.sqrt(l)
    Q = np.sum(Xs * Xs, axis=-1, keepdims=True)
    dist = Q + np.swapaxes(Q, -1, -2) - 2.0 * np.matmul(Xs, np.swapaxes(Xs, -1, -2))
    return sigma * np.exp(-0.5 * dist)


def computeNKmm(X, l, sigma, sigma0):
    """utils.py:765-783 -- noise on the whole diagonal, no jitter."""
    return computeKmm(X, l, sigma) + np.eye(X.shape[0]) * sigma0


def chol2inv(mat):
    """utils.py:657-674: L=chol(A); L^-1 by a general solve against I; A^-1 = L^-T L^-1."""
    L = np.linalg.cholesky(mat)
    invlower = np.linalg.solve(L, np.eye(mat.shape[-1]))
    return np.swapaxes(invlower, -1, -2) @ invlower


def multichol2inv(mats):
    """utils.py:677-700 -- batched chol2inv."""
    return np.stack([chol2inv(m) for m in mats])


def precomputeInvKs(ls, sigmas, sigma0s, Xsamples):
    """utils.py:1838-1856."""
    return np.stack(
        [chol2inv(computeNKmm(Xsamples, ls[i], sigmas[i], sigma0s[i])) for i in range(len(sigmas))]
    )


def eval_invKmaxsams(ls, sigmas, sigma0s, Xsamples, maxima):
    """utils.py:1859-1883 -- maximizer-conditioned (n+1)x(n+1) inverses, one per maximizer.

    NOTE the reference calls computeNKmm on [x*_j; X], i.e. noise on ALL n+1 diagonal entries.
    """
    out = []
    for i in range(len(sigmas)):
        row = []
        for j in range(maxima.shape[1]):
            xm = np.concatenate([maxima[i, j].reshape(1, -1), Xsamples], axis=0)
            row.append(chol2inv(computeNKmm(xm, ls[i], sigmas[i], sigma0s[i])))
        out.append(np.stack(row))
    return np.stack(out)


def get_pNK_test_obs(ls, sigmas, sigma0s, nhyp, X, test_xs):
    """evaluate_mp.py:12-53 (5 identical copies): K([test;X]) + sigma0*diag(0_T, I_n); LU inverse."""
    nobs = X.shape[0]
    ntest = test_xs.shape[0]
    pNKs, invs = [], []
    for i in range(nhyp):
        Xto = np.concatenate([test_xs, X], axis=0)
        K = computeKmm(Xto, ls[i], sigmas[i])
        noise = np.zeros((ntest + nobs, ntest + nobs))
        noise[ntest:, ntest:] = np.eye(nobs) * sigma0s[i]
        pNK = K + noise
        pNKs.append(pNK)
        invs.append(np.linalg.inv(pNK))
    return np.stack(pNKs), np.stack(invs)


def compute_mean_var_f(x, Xsamples, Ysamples, l, sigma, sigma0, NKInv=None, fullcov=False):
    """utils.py:786-815 -- GP posterior mean / (diag | full) covariance, diag clipped at 1e-100."""
    if NKInv is None:
        NKInv = chol2inv(computeNKmm(Xsamples, l, sigma, sigma0))
    kstar = computeKnm(x, Xsamples, l, sigma)
    mean = np.squeeze(kstar @ (NKInv @ Ysamples.reshape(-1, 1)))
    if fullcov:
        Kx = computeKmm(x, l, sigma)
        var = Kx - kstar @ NKInv @ kstar.T
        d = np.clip(np.diag(var), CLIP_MIN, np.inf)
        var = var.copy()
        np.fill_diagonal(var, d)
    else:
        var = sigma - np.sum((kstar @ NKInv) * kstar, axis=1)
        var = np.clip(var, CLIP_MIN, np.inf)
    return mean, var


def compute_mean_f_np(x, Xsamples, Ysamples, l, sigma, sigma0):
    """utils.py:862-885 -- the function behind the Brooms-Barn objectives (functions.py:383-387)."""
    m = Xsamples.shape[0]
    NK = computeKmm(Xsamples, l, sigma) + np.eye(m) * sigma0
    inv = np.linalg.inv(NK)
    kstar = computeKnm(x.reshape(-1, Xsamples.shape[1]), Xsamples, l, sigma)
    return kstar.dot(inv.dot(Ysamples.reshape(m, 1))).reshape(-1)


def evaluate_norm_entropy(s):
    """utils.py:653-654 (float32-rounded constant, Q12)."""
    return NORM_ENTROPY_CONST + np.log(s) + 0.5


# --------------------------------------------------------------------------------------
# TFP / TF primitives
# --------------------------------------------------------------------------------------
def normal_log_prob(x, loc, scale):
    z = (x - loc) / scale
    return -0.5 * z * z - (HALF_LOG_2PI + np.log(scale))


def reduce_logsumexp(x, axis):
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        raw_max = np.max(x, axis=axis, keepdims=True)
        my_max = np.where(np.isfinite(raw_max), raw_max, 0.0)
        out = np.log(np.sum(np.exp(x - my_max), axis=axis, keepdims=True)) + my_max
    return np.squeeze(out, axis=axis)


# --------------------------------------------------------------------------------------
# Stage A: random-Fourier-feature function samples and their arg-max
# --------------------------------------------------------------------------------------
def draw_random_init_weights_features(xdim, n_funcs, n_features, xx, yy, l, sigma, sigma0,
                                      randn_W, rand_b, randn_noise):
    """optfunc.py:303-385 with the three global-RNG draws made explicit, in the reference's
    draw order: randn(S,F,d) -> W (:329); rand(S,F,1) -> b (:334); randn(S,F,1) -> noise (:344)."""
    n = xx.shape[0]
    l = np.asarray(l, dtype=np.float64).reshape(1, xdim)
    xx_t = np.tile(xx.reshape(1, n, xdim), (n_funcs, 1, 1))
    yy_t = np.tile(yy.reshape(1, n, 1), (n_funcs, 1, 1))
    idn = np.tile(np.eye(n).reshape(1, n, n), (n_funcs, 1, 1))
    W = randn_W * np.tile(np.sqrt(l).reshape(1, 1, xdim), (n_funcs, n_features, 1))
    b = 2.0 * np.pi * rand_b
    Z = np.sqrt(2.0 * sigma / n_features) * np.cos(
        np.matmul(W, np.transpose(xx_t, (0, 2, 1))) + np.tile(b, (1, 1, n)))
    noise = randn_noise
    if n < n_features:
        Sigma = np.matmul(np.transpose(Z, (0, 2, 1)), Z) + sigma0 * idn
        mu = np.matmul(np.matmul(Z, np.linalg.inv(Sigma)), yy_t)
        e, v = np.linalg.eig(Sigma)
        e = e.reshape(n_funcs, n, 1)
        r = 1.0 / (np.sqrt(e) * (np.sqrt(e) + np.sqrt(sigma0)))
        theta = noise - np.matmul(
            Z, np.matmul(v, r * np.matmul(np.transpose(v, (0, 2, 1)),
                                          np.matmul(np.transpose(Z, (0, 2, 1)), noise)))) + mu
    else:
        Sigma = np.linalg.inv(
            np.matmul(Z, np.transpose(Z, (0, 2, 1))) / sigma0
            + np.tile(np.eye(n_features).reshape(1, n_features, n_features), (n_funcs, 1, 1)))
        mu = np.matmul(np.matmul(Sigma, Z), yy_t) / sigma0
        theta = mu + np.matmul(np.linalg.cholesky(Sigma), noise)
    return theta, W, b


def rff_fvals(x, theta, W, b, sigma):
    """f_j(x) = sqrt(2 sigma/F) theta_j^T cos(W_j x^T + b_j) for all samples j.

    utils_for_continuous.py:172-180 (TF) / optfunc.py:389-402 (numpy twin).
    theta (S,F,1), W (S,F,d) or (1,F,d) [shared], b (S,F,1) or (1,F,1) -> (S, N).
    """
    S, F = theta.shape[0], theta.shape[1]
    out = np.empty((S, x.shape[0]))
    for j in range(S):
        Wj = W[j if W.shape[0] > 1 else 0]
        bj = b[j if b.shape[0] > 1 else 0]
        out[j] = np.squeeze(
            np.sqrt(2.0 * sigma / F) * np.matmul(theta[j].T, np.cos(np.matmul(Wj, x.T) + bj)))
    return out


def adam_refine_maximizers(x0, theta, W, b, sigma, xmin, xmax, ntrain, lr=1e-3, beta1=0.9, beta2=0.999,
                           epsilon=1e-8):
    """utils_for_continuous.py:201-232: TF-1 AdamOptimizer() on -sum_{j,i} f_j(xs[j,i]) for `ntrain` steps with the
    clip_by_value(xmin, xmax) variable constraint applied after every step.  x0 (S,k,d); theta (S,F,1);
    W (S,F,d) or (1,F,d); b (S,F,1) or (1,F,1).  Returns (xs (S,k,d), fvals (S,k)).
    TF-1 Adam (tensorflow/python/training/adam.py): lr_t = lr sqrt(1-beta2^t)/(1-beta1^t);
    m <- beta1 m + (1-beta1) g; v <- beta2 v + (1-beta2) g^2; var <- var - lr_t m / (sqrt(v) + epsilon).
    Pinned on the reference's own graph code run on oracle/tf_shim/graph.py (tests/golden/adam.npz)."""
    S, k, d = x0.shape
    F = theta.shape[1]
    scale = np.sqrt(2.0 * sigma / F)
    th = theta.reshape(S, F)
    xs = x0.copy()
    m = np.zeros_like(xs)
    v = np.zeros_like(xs)
    lo = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (d,))
    hi = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (d,))

    def args(xs_):
        Wb = np.broadcast_to(W, (S, F, d))
        return np.einsum("jfc,jic->jif", Wb, xs_) + np.broadcast_to(b.reshape(-1, 1, F), (S, k, F))

    for t in range(1, ntrain + 1):
        h = args(xs)                                                        # (S,k,F)
        g = scale * np.einsum("jif,jfc->jic", th[:, None, :] * np.sin(h), np.broadcast_to(W, (S, F, d)))
        lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
        m = beta1 * m + (1.0 - beta1) * g
        v = beta2 * v + (1.0 - beta2) * g * g
        xs = np.clip(xs - lr_t * m / (np.sqrt(v) + epsilon), lo, hi)
    fvals = scale * np.sum(th[:, None, :] * np.cos(args(xs)), axis=2)
    return xs, fvals


def top_k_lowest_index(v, k):
    """tf.math.top_k semantics: descending value, ties -> lower index first."""
    order = np.lexsort((np.arange(v.shape[0]), -v))
    return order[:k]


def rff_topk(x, theta, W, b, sigma, k):
    """utils_for_continuous.py:183 -- indices (S,k) and values (S,k) of the top-k candidates."""
    f = rff_fvals(x, theta, W, b, sigma)
    idx = np.stack([top_k_lowest_index(f[j], k) for j in range(f.shape[0])])
    val = np.take_along_axis(f, idx, axis=1)
    return idx, val


def discrete_maximizer_samples(xs, X, Y, l, sigma, sigma0, z):
    """bo_discrete.py:359-400 + utils.sqrtm (utils.py:49-53): exact joint posterior sample on the
    grid, f = sqrtm(Sigma) z^T + mu with sqrtm by SVD (U sqrt(S) V^T), arg-max over the grid.
    z (nmax, nx) standard normal.  Returns (max_idx (nmax,), max_f (nmax,))."""
    mean, cov = compute_mean_var_f(xs, X, Y, l, sigma, sigma0, fullcov=True)
    U, s, Vt = np.linalg.svd(cov)
    sq = U @ np.diag(np.sqrt(s)) @ Vt
    f = sq @ z.T + mean.reshape(-1, 1)
    return np.argmax(f, axis=0), np.max(f, axis=0)


# --------------------------------------------------------------------------------------
# Stage C: trusted-maximizer statistics on the host  (utils.py:1554-1833, ep.py,
# empirical_approximation.py)
# --------------------------------------------------------------------------------------
def get_duplicate_mask(xs, resolution=1e-5):
    """utils.py:1554-1573."""
    n = xs.shape[0]
    mask = np.zeros(n)
    for i in range(n):
        if mask[i] == 1.0:
            continue
        for j in range(i + 1, n):
            d = xs[j] - xs[i]
            if np.sqrt(np.sum(d * d)) <= resolution:
                mask[j] = 1.0
    return mask


def remove_duplicates(xs, resolution=1e-5):
    """utils.py:1576-1581."""
    return np.delete(xs, np.where(get_duplicate_mask(xs, resolution) == 1.0)[0], axis=0)


def mvn_transform_matrix(cov):
    """utils.py:1668-1670 -- A = V diag(sqrt(clip(e,0))) from a GENERAL eig of the symmetric cov.
    The columns' order and signs are LAPACK dgeev's; callers that need identical samples for
    identical z must use this same matrix (it is an explicit input of the device path)."""
    e, v = np.linalg.eig(cov)
    e = np.clip(e, 0.0, np.inf)
    return v.dot(np.diag(np.sqrt(e)))


def sample_multivariate_normal_maxidx(mean, cov, nsample, n_min_sample=1, z=None,
                                      remove_correlated_dims=True, transform=None):
    """utils.py:1654-1729 with z = np.random.normal(size=(nsample,n,1)) (:1674) explicit."""
    n = cov.shape[0]
    mean = mean.reshape(n)
    A = mvn_transform_matrix(cov) if transform is None else transform
    # NOTE (reference quirk Q11): mean.reshape(1,n) broadcasts against the (nsample,n,1) product to
    # (nsample,n,n) and the mean over the LAST axis then adds average(mean) -- not mean[i] -- to every
    # coordinate.  Restated literally.
    samples = np.mean(mean.reshape(1, n) + np.matmul(A[None], z), axis=-1)  # (nsample, n)
    corr_idxs = []
    if remove_correlated_dims:
        with np.errstate(invalid="ignore", divide="ignore"):
            corr = np.tril(np.corrcoef(samples.T), k=-1)
        rows, _ = np.where(corr > 1 - 1e-6)
        corr_idxs = rows
    ign = samples.copy()
    ign[:, corr_idxs] = -np.inf
    maxidxs = np.argmax(ign, axis=1)
    unique, groups, sizes = [], [], []
    for i in range(n):
        idxs = np.where(maxidxs == i)[0]
        if len(idxs) >= n_min_sample:
            groups.append(samples[idxs, :])
            sizes.append(len(idxs))
            unique.append(i)
    nmax = len(unique)
    unique = np.array(unique)
    sizes = np.array(sizes)
    K = int(np.max(sizes))
    ret = np.zeros([nmax, n, K])
    masks = np.zeros([nmax, K]).astype(int)
    for i in range(nmax):
        ret[i, :, :sizes[i]] = groups[i].T
        ret[i, :, sizes[i]:] = np.tile(mean.reshape(n, 1), (1, K - sizes[i]))  # prior mean pad
        masks[i, :sizes[i]] = 1
    return nmax, unique, K, ret, masks, sizes


def get_empirical_stat_from_samples(samples, weight=None):
    """empirical_approximation.py:86-107."""
    n = samples.shape[0]
    if weight is None:
        weight = np.ones(n)
    tw = np.sum(weight)
    weight = weight.reshape(-1, 1)
    emean = np.sum(samples * weight, axis=0, keepdims=True) / tw
    zs = (samples - emean) * np.sqrt(weight)
    ecov = zs.T.dot(zs) / (tw - 1.0)
    return emean.reshape(-1, 1), ecov


def _ep_approx(cs, mean_xs, inv_cov_xs, site_means, site_vars):
    """ep.py:8-28."""
    csp = cs / np.sqrt(site_vars.reshape(1, -1))
    approx_cov = np.linalg.inv(inv_cov_xs + csp.dot(csp.T))
    approx_mean = approx_cov.dot(
        inv_cov_xs.dot(mean_xs)
        + np.sum(cs * site_means.reshape(1, -1) / site_vars.reshape(1, -1), axis=1, keepdims=True))
    return approx_mean, approx_cov


def approximate_EP(max_idx, mean_xs, cov_xs, resolution=1e-9, max_niter=1000, return_niter=False):
    """ep.py:73-204 -- EP for N(mu,Sigma) truncated to {f_max_idx >= f_i for all i}.

    One site is updated per outer iteration (index count_niter % (nx-1), :162), invalid site
    variances are skipped (:167-172), the approximation is recomputed by a full inverse (:182).
    """
    with np.errstate(all="ignore"):
        mean_xs = mean_xs.reshape(-1, 1)
        nx = mean_xs.shape[0]
        cs = np.zeros([nx, nx - 1])
        j = 0
        for i in range(nx):
            if i != max_idx:
                cs[max_idx][j] = 1
                cs[i][j] = -1
                j += 1
        site_means = cs.T.dot(mean_xs).reshape(nx - 1)
        site_vars = np.diag(cs.T.dot(cov_xs).dot(cs)).copy()
        inv_cov_xs = np.linalg.inv(cov_xs)
        approx_mean, approx_cov = _ep_approx(cs, mean_xs, inv_cov_xs, site_means, site_vars)
        count = 0
        diff = 1e10
        while diff > resolution and count < max_niter:
            count += 1
            # cavity, ep.py:31-49
            ctSc = np.diag(cs.T.dot(approx_cov).dot(cs))
            cav_vars = 1.0 / (1.0 / ctSc - 1.0 / site_vars)
            cav_means = cav_vars * (cs.T.dot(approx_mean).squeeze() / ctSc - site_means / site_vars)
            # truncated-normal moments, ep.py:52-60
            beta = cav_means / np.sqrt(cav_vars)
            poc = spst.norm.pdf(beta) / spst.norm.cdf(beta)
            cf_means = cav_means + np.sqrt(cav_vars) * poc
            cf_vars = (-cav_means * np.sqrt(cav_vars) * poc - cf_means * cf_means
                       + 2.0 * cav_means * cf_means - cav_means * cav_means + cav_vars)
            # site update, ep.py:63-70
            upd_vars = 1.0 / (1.0 / cf_vars - 1.0 / cav_vars)
            upd_means = upd_vars * (cf_means / cf_vars - cav_means / cav_vars)
            uidx = count % (nx - 1)
            last = count - 1 if count >= 1 else nx - 2
            while (upd_vars[uidx] <= 0 or np.isnan(upd_vars[uidx]) or np.isinf(upd_vars[uidx])) \
                    and uidx != last:
                count += 1
                uidx = count % (nx - 1)
            site_means[uidx] = upd_means[uidx]
            site_vars[uidx] = upd_vars[uidx]
            new_mean, new_cov = _ep_approx(cs, mean_xs, inv_cov_xs, site_means, site_vars)
            dm = new_mean - approx_mean
            dc = new_cov - approx_cov
            diff = np.mean(dm * dm + dc * dc)
            if np.any(np.isnan(new_cov)) or np.any(np.isnan(new_mean)):
                break
            approx_mean, approx_cov = new_mean, new_cov
    if return_niter:
        return approx_mean, approx_cov, count
    return approx_mean, approx_cov


def get_testidxs_stats(nhyp, test_xs, test_means, test_covs, mode="sample", nsample=1000,
                       n_min_sample=2, z=None, transforms=None):
    """utils.py:1733-1833.  z: (nhyp, nsample, ntest, 1) -- the draws of :1674, one per hyp.
    mode='ep' calls approximate_EP with the arguments utils.py:1820-1825 intends (the
    shipped code raises NameError there because utils.py never imports ep)."""
    ntest = test_xs.shape[0]
    maxidxs = set(range(ntest))
    max_probs = np.ones([nhyp, ntest]) / ntest
    K_all = 0
    post_samples = np.zeros([nhyp, ntest, ntest, nsample])
    post_masks = np.zeros([nhyp, ntest, nsample]).astype(bool)
    for i in range(nhyp):
        nmax, uniq, K, rs, rm, sizes = sample_multivariate_normal_maxidx(
            test_means[i, :], test_covs[i, :, :], nsample, n_min_sample, z=z[i],
            transform=None if transforms is None else transforms[i])
        K_all = max(K_all, K)
        maxidxs = maxidxs.intersection(uniq)
        max_probs[i, uniq] = sizes / np.sum(sizes)
        for j, mi in enumerate(uniq):
            post_samples[i, mi, :, :K] = rs[j]
            post_masks[i, mi, :K] = rm[j]
    remove = np.array(sorted(set(range(ntest)).difference(maxidxs)), dtype=int)
    if len(remove) != 0:
        test_xs = np.delete(test_xs, remove, axis=0)
        test_means = np.delete(test_means, remove, axis=1)
        test_covs = np.delete(np.delete(test_covs, remove, axis=1), remove, axis=2)
        max_probs = np.delete(max_probs, remove, axis=1)
    max_probs = max_probs / np.sum(max_probs, axis=1, keepdims=True)
    ntest = test_xs.shape[0]
    if len(remove) != 0:
        post_samples = np.delete(np.delete(post_samples, remove, axis=1), remove, axis=2)
        post_masks = np.delete(post_masks, remove, axis=1)
    post_samples = post_samples[..., :K_all]
    post_masks = post_masks[..., :K_all]
    if mode == "sample":
        return test_xs, max_probs, test_means, test_covs, post_samples, post_masks
    pm = np.zeros([nhyp, ntest, ntest])
    pc = np.zeros([nhyp, ntest, ntest, ntest])
    for i in range(nhyp):
        for j in range(ntest):
            if mode == "empirical":
                m, c = get_empirical_stat_from_samples(
                    post_samples[i, j].T, weight=post_masks[i, j].astype(int))
            elif mode == "ep":
                m, c = approximate_EP(j, test_means[i].reshape(-1, 1), test_covs[i],
                                      resolution=1e-9, max_niter=100)
            else:
                raise Exception("Unknown mode in get_testidxs_stats!")
            pm[i, j] = m.squeeze()
            pc[i, j] = c
    return test_xs, max_probs, test_means, test_covs, pm, pc


# --------------------------------------------------------------------------------------
# Stage B/D: the TES criteria
# --------------------------------------------------------------------------------------
def _gather_nonzero(max_probs, *arrs):
    idx = np.where(max_probs > 0.0)[0]
    return (max_probs[idx],) + tuple(a[idx] for a in arrs)


def get_queried_f_stat_given_data_maxidx(xs, l, sigma, X, Y, test_xs, invpNK, post_mean, post_cov):
    """evaluate_mp.py:56-113 -> (query_mean (nmax,nx), query_var (nmax,nx))."""
    ntest = test_xs.shape[0]
    nobs = X.shape[0]
    k = computeKnm(xs, np.concatenate([test_xs, X], axis=0), l, sigma)
    Kq = sigma - np.sum((k @ invpNK) * k, axis=1)
    M = k @ invpNK[:, :ntest]
    b = k @ invpNK[:, ntest:] @ Y.reshape(nobs, 1)
    qmean = (M @ post_mean.T + b).T
    qvar = Kq.reshape(1, -1) + np.sum((M[None] @ post_cov) * M[None], axis=2)
    return qmean, qvar


def evaluate_mp(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs,
                max_probs_all, post_mean_all, post_cov_all, invKobs_all, invpNK_all,
                stochastic=False, niteration=10, z=None):
    """TES_ep, q=1: evaluate_mp.py:116-295.
    z: (nhyp, niteration, nysample, nmax', nx) standard normals (the draws of :265), where nmax'
    counts the maximizers with max_probs>0 of that hyp."""
    avg = np.zeros(nx)
    for i in range(nhyp):
        p, pm, pc = _gather_nonzero(max_probs_all[i], post_mean_all[i], post_cov_all[i])
        qmean, qvar = get_queried_f_stat_given_data_maxidx(
            x, ls[i], sigmas[i], X, Y, test_xs, invpNK_all[i], pm, pc)
        with np.errstate(invalid="ignore"):
            qstd = np.sqrt(qvar + sigma0s[i])  # no clipping
        if not stochastic:
            _, varf = compute_mean_var_f(x, X, Y, ls[i], sigmas[i], sigma0s[i], invKobs_all[i])
            ent_y = evaluate_norm_entropy(np.sqrt(varf.reshape(nx) + sigma0s[i]))
            cond = evaluate_norm_entropy(qstd)
            val = ent_y - np.mean(cond * p[:, None], axis=0)  # reduce_MEAN over j
        else:
            s = np.zeros(nx)
            for it in range(niteration):
                s += _mp_each_batch_y_sample(p, qmean, qstd, z[i][it])
            val = s / niteration
        avg = avg + val / nhyp
    return avg


def _mp_each_batch_y_sample(p, qmean, qstd, z):
    """evaluate_mp.py:248-295.  z (nysample, nmax, nx)."""
    nmax = p.shape[0]
    y = qmean[None] + qstd[None] * z
    lp = normal_log_prob(y, qmean[None], qstd[None])
    cond = -np.mean(np.sum(lp * p.reshape(1, nmax, 1), axis=1), axis=0)
    # marginal[iy, j, j', x] = y[iy, j, x] scored under component j'
    lpm = normal_log_prob(y[:, :, None, :], qmean[None, None], qstd[None, None])
    with np.errstate(divide="ignore"):
        lpm = lpm + np.log(p.reshape(1, 1, nmax, 1))
    lm = reduce_logsumexp(lpm, axis=2)
    ent = -np.mean(np.sum(lm * p.reshape(1, nmax, 1), axis=1), axis=0)
    return ent - cond


def evaluate_mp_lite(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, test_xs, max_probs_all,
                     post_mean_all, post_cov_all, invpNK_all):
    """evaluate_mp_lite.py:116-225: sum_ab p_a p_b KL-like term, v = var + sigma0."""
    avg = np.zeros(nx)
    for i in range(nhyp):
        p, pm, pc = _gather_nonzero(max_probs_all[i], post_mean_all[i], post_cov_all[i])
        qmean, qvar = get_queried_f_stat_given_data_maxidx(
            x, ls[i], sigmas[i], X, Y, test_xs, invpNK_all[i], pm, pc)
        v = qvar + sigma0s[i]
        m0, m1 = qmean[None, :, :], qmean[:, None, :]
        v0, v1 = v[None, :, :], v[:, None, :]
        diff = m0 - m1
        with np.errstate(invalid="ignore", divide="ignore"):
            mips = 0.5 * np.log(v1 / v0) + 0.5 * (v0 + diff * diff) / v1 - 0.5
        p0, p1 = p.reshape(1, -1, 1), p.reshape(-1, 1, 1)
        avg = avg + np.sum(np.sum(p0 * p1 * mips, axis=0), axis=0) / nhyp
    return avg


def get_queried_f_stat_given_test_samples(x, l, sigma, X, Y, test_xs, invpNK_test, invpNK_obs,
                                          post_test_samples):
    """evaluate_sample_mp.py:54-129 -> (query_mean (nmax,nx,nsample), query_var (nx,))."""
    nobs = X.shape[0]
    k = computeKnm(x, np.concatenate([test_xs, X], axis=0), l, sigma)
    tmp_test = k @ invpNK_test  # (nx, ntest)
    qm_test = np.einsum("xt,jts->jxs", tmp_test, post_test_samples)
    qm_obs = k @ (invpNK_obs @ Y.reshape(nobs, 1))  # (nx,1)
    qmean = qm_test + qm_obs[None]
    tmp = k @ np.concatenate([invpNK_test, invpNK_obs], axis=1)
    qvar = sigma - np.sum(tmp * k, axis=1)
    return qmean, qvar


def evaluate_sample_mp(x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs,
                       max_probs_all, post_test_samples_all, post_test_mask_all, invpNK_all,
                       niteration=10, z=None):
    """TES_sp, q=1: evaluate_sample_mp.py:132-336.
    z: (nhyp, niteration, nysample, nmax', nx, nsample) (the draws of :268)."""
    ntest = post_test_samples_all.shape[2]
    avg = np.zeros(nx)
    for i in range(nhyp):
        p, ps, msk = _gather_nonzero(max_probs_all[i], post_test_samples_all[i],
                                     post_test_mask_all[i].astype(bool))
        inv = invpNK_all[i]
        qmean, qvar = get_queried_f_stat_given_test_samples(
            x, ls[i], sigmas[i], X, Y, test_xs, inv[:, :ntest], inv[:, ntest:ntest + nobs], ps)
        with np.errstate(invalid="ignore"):
            qstd = np.sqrt(qvar + sigma0s[i]).reshape(1, nx, 1)
        s = np.zeros(nx)
        for it in range(niteration):
            s += _sample_mp_each(p, qmean, qstd, msk, z[i][it])
        avg = avg + (s / niteration) / nhyp
    return avg


def _sample_mp_each(p, qmean, qstd, msk, z):
    """evaluate_sample_mp.py:250-336.  qmean (nmax,nx,K), qstd (1,nx,1), msk (nmax,K),
    z (nysample,nmax,nx,K)."""
    nmax, nx, K = qmean.shape
    y = qmean[None] + qstd[None] * z
    lp = normal_log_prob(y, qmean[None], qstd[None])
    m4 = msk.reshape(1, nmax, 1, K)
    lp = np.where(m4, lp, -np.inf)
    lmix = reduce_logsumexp(lp, axis=3)  # un-normalised over K
    with np.errstate(invalid="ignore"):
        cond = -np.mean(np.sum(lmix * p.reshape(1, nmax, 1), axis=1), axis=0)
    # marginal[iy, j, j', x, s] = y[iy, j, x, s] scored under (j', same s); mask of SOURCE j (Q3)
    lpm = normal_log_prob(y[:, :, None, :, :], qmean[None, None], qstd[None, None])
    lpm = np.where(m4[:, :, None, :, :], lpm, -np.inf)
    lmm = reduce_logsumexp(lpm, axis=4)  # (Y, j, j', x)
    with np.errstate(divide="ignore"):
        lmm = lmm + np.log(p.reshape(1, 1, nmax, 1))
    lmarg = reduce_logsumexp(lmm, axis=2)
    with np.errstate(invalid="ignore"):
        ent = -np.mean(np.sum(lmarg * p.reshape(1, nmax, 1), axis=1), axis=0)
    return ent - cond


def get_batch_queried_f_stat_given_test_samples(x, l, sigma, X, Y, test_xs, invpNK_test,
                                                invpNK_obs, post_test_samples):
    """evaluate_batch_sample_mp.py:54-157 -> mean (nmax,nx,q,K), cov (nx,q,q)."""
    nx, q, d = x.shape
    nobs = X.shape[0]
    k = computeKnm(x.reshape(-1, d), np.concatenate([test_xs, X], axis=0), l, sigma)
    k = k.reshape(nx, q, -1)
    k_x = computeKmm(x, l, sigma)
    qm_obs = k @ invpNK_obs @ Y.reshape(nobs, 1)  # (nx,q,1)
    qm_test = np.einsum("xqm,mt,jts->jxqs", k, invpNK_test, post_test_samples)
    qmean = qm_obs[None] + qm_test
    inv = np.concatenate([invpNK_test, invpNK_obs], axis=1)
    qcov = k_x - k @ inv @ np.transpose(k, (0, 2, 1))
    return qmean, qcov


def evaluate_batch_sample_mp(x, ls, sigmas, sigma0s, X, Y, xdim, nobs, nhyp, nysample, test_xs,
                             max_probs_all, post_test_samples_all, post_test_mask_all, invpNK_all,
                             niteration=10, z=None):
    """TES_sp, batch q: evaluate_batch_sample_mp.py:160-382.  x (nx,q,d).
    z: (nhyp, niteration, nysample, nmax', nx, nsample, q) (the draws of :315)."""
    nx, q, _ = x.shape
    ntest = post_test_samples_all.shape[2]
    avg = np.zeros(nx)
    for i in range(nhyp):
        p, ps, msk = _gather_nonzero(max_probs_all[i], post_test_samples_all[i],
                                     post_test_mask_all[i].astype(bool))
        inv = invpNK_all[i]
        qmean, qcov = get_batch_queried_f_stat_given_test_samples(
            x, ls[i], sigmas[i], X, Y, test_xs, inv[:, :ntest], inv[:, ntest:ntest + nobs], ps)
        qcov = qcov + np.eye(q)[None] * sigma0s[i]
        qmean = np.transpose(qmean, (0, 1, 3, 2))  # (nmax,nx,K,q)
        L = np.linalg.cholesky(qcov)  # (nx,q,q)
        s = np.zeros(nx)
        for it in range(niteration):
            s += _batch_sample_mp_each(p, qmean, L, msk, z[i][it])
        avg = avg + (s / niteration) / nhyp
    return avg


def _mvn_tril_log_prob(y, loc, L):
    """log N(y; loc, L L^T); y,loc (..., nx, K, q) broadcastable; L (nx,q,q)."""
    q = L.shape[-1]
    diff = y - loc
    # solve L w = diff  per candidate
    Linv = np.linalg.inv(L)  # small q; (nx,q,q)
    w = np.einsum("xab,...xsb->...xsa", Linv, diff)
    logdet = np.sum(np.log(np.diagonal(L, axis1=-2, axis2=-1)), axis=-1)  # (nx,)
    return -0.5 * np.sum(w * w, axis=-1) - logdet[:, None] - 0.5 * q * math.log(2.0 * math.pi)


def _batch_sample_mp_each(p, qmean, L, msk, z):
    """evaluate_batch_sample_mp.py:294-382.  qmean (nmax,nx,K,q), L (nx,q,q), msk (nmax,K),
    z (nysample,nmax,nx,K,q)."""
    nmax, nx, K, q = qmean.shape
    y = qmean[None] + np.einsum("xab,yjxsb->yjxsa", L, z)
    lp = _mvn_tril_log_prob(y, qmean[None], L)  # (Y,nmax,nx,K)
    m4 = msk.reshape(1, nmax, 1, K)
    lp = np.where(m4, lp, -np.inf)
    lmix = reduce_logsumexp(lp, axis=3)
    with np.errstate(invalid="ignore"):
        cond = -np.mean(np.sum(lmix * p.reshape(1, nmax, 1), axis=1), axis=0)
    lpm = _mvn_tril_log_prob(y[:, :, None], qmean[None, None], L)  # (Y,j,j',nx,K)
    lpm = np.where(m4[:, :, None, :, :], lpm, -np.inf)
    lmm = reduce_logsumexp(lpm, axis=4)
    with np.errstate(divide="ignore"):
        lmm = lmm + np.log(p.reshape(1, 1, nmax, 1))
    lmarg = reduce_logsumexp(lmm, axis=2)
    with np.errstate(invalid="ignore"):
        ent = -np.mean(np.sum(lmarg * p.reshape(1, nmax, 1), axis=1), axis=0)
    return ent - cond


def get_batch_queried_f_stat_given_data_maxidx(xs, l, sigma, X, Y, test_xs, invpNK, post_mean,
                                               post_cov):
    """evaluate_batch_mp.py:60-142 -> mean (nmax,nx,q), cov (nmax,nx,q,q)."""
    nx, q, d = xs.shape
    ntest = test_xs.shape[0]
    nobs = X.shape[0]
    k = computeKnm(xs.reshape(-1, d), np.concatenate([test_xs, X], axis=0), l, sigma)
    k = k.reshape(nx, q, -1)
    k_x = computeKmm(xs, l, sigma)
    Kq = k_x - (k @ invpNK) @ np.transpose(k, (0, 2, 1))
    M = k @ invpNK[:, :ntest]  # (nx,q,ntest)
    b = k @ (invpNK[:, ntest:] @ Y.reshape(nobs, 1))  # (nx,q,1)
    qmean = np.transpose(M @ post_mean.T[None] + b, (2, 0, 1))
    qvar = Kq[None] + M[None] @ post_cov[:, None] @ np.transpose(M, (0, 2, 1))[None]
    return qmean, qvar


def evaluate_batch_mp(x, ls, sigmas, sigma0s, X, Y, xdim, nobs, nhyp, nysample, test_xs,
                      max_probs_all, post_mean_all, post_cov_all, invKobs_all, invpNK_all,
                      niteration=10, z=None):
    """TES_ep, batch q: evaluate_batch_mp.py:145-331.  As shipped this evaluates to
    -(sum p) log(sum p) ~ 0 for every input: the log-prob is taken under the
    STANDARD normal (:302) and the marginal is built by tiling that same log-prob (:317-322).
    z: (nhyp, niteration, nysample, nmax', nx, q)."""
    nx, q, _ = x.shape
    avg = np.zeros(nx)
    for i in range(nhyp):
        p, pm, pc = _gather_nonzero(max_probs_all[i], post_mean_all[i], post_cov_all[i])
        nmax = p.shape[0]
        qmean, qcov = get_batch_queried_f_stat_given_data_maxidx(
            x, ls[i], sigmas[i], X, Y, test_xs, invpNK_all[i], pm, pc)
        qcov = qcov + (np.eye(q) * sigma0s[i])[None, None]
        ev, evec = np.linalg.eigh(qcov)
        ev = np.clip(ev, 0.0, np.inf)
        T = evec @ (np.sqrt(ev)[..., None] * np.eye(q))  # evec @ diag(sqrt(ev))
        s = np.zeros(nx)
        for it in range(niteration):
            zz = z[i][it]  # (Y,nmax,nx,q)
            y = np.mean(T[None] @ zz[..., None], axis=-1) + qmean[None]
            lp = -0.5 * np.sum(y * y, axis=-1) - 0.5 * q * math.log(2.0 * math.pi)  # (Y,nmax,nx)
            cond = -np.mean(np.sum(lp * p.reshape(1, nmax, 1), axis=1), axis=0)
            lpm = np.tile(lp[:, :, None, :], (1, 1, nmax, 1)) + np.log(p.reshape(1, 1, nmax, 1))
            lm = reduce_logsumexp(lpm, axis=2)
            ent = -np.mean(np.sum(lm * p.reshape(1, nmax, 1), axis=1), axis=0)
            s += ent - cond
        avg = avg + (s / niteration) / nhyp
    return avg


def optimize_criterion(vals):
    """bo_discrete.py:794-807 / :1053-1055: first maximum wins."""
    idx = int(np.argmax(vals))
    return idx, vals[idx]
