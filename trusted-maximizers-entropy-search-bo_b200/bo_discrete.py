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


def run(f, xs, l, sigma, sigma0, criterion="ftl", mode="ep", nquery=50, nrun=1, ninit=2, nmax=5, nhyp=1,
        deterministic=0, nsto=50, nysample=10, ntestsample=1000, seed=1, duplicate_resolution=1e-3,
        sampler="exact", n_features=300, transform_fn=None, verbose=False):
    """The BO loop of bo_discrete.py:928-1145 for the TES criteria, re-hosted on the device kernels (no TF session).

    f: objective on (n, xdim) arrays; xs: the discrete domain (nx, xdim) = xs_to_optimize; (l, sigma, sigma0): the
    known GP hyper-parameters (bo_discrete.py:835-840; `l` inverse-squared length-scales); the other arguments are the
    driver's flags (`-c criterion`, `--mode`, `-q numqueries`, `-r numruns`, `--ninit`, `-m nmax`, `-d deterministic`,
    `-o nsto`, `-y nysample`) and constants (`ntestsample` = 1000 :881, `max_n_test` = nmax :880, duplicate_resolution
    :189, seed :842).  Per run the host RNG is seeded with seed + run (:933-935) and consumed in the reference's order:
    initial design `randint` + observation noise `randn` (:937-941); per query the test-set top-up `shuffle` (:581-583),
    the joint-sample draw inside utils.get_testidxs_stats (utils.py:1674) and the observation noise (:1059).
    Two draws live inside the TF graph in the reference and cannot be replayed from numpy: the (nmax, nx) normals of
    the exact discrete sampler (:359-362) -- drawn here with np.random.randn at the top of the query -- and the
    y samples of the stochastic criteria (counter-based, keyed on seed + run and the query number).
    sampler='exact': the O(nx^3) joint posterior sample of :359-400; 'rff': stage A (optfunc draw with `n_features`
    features + per-sample arg-max over the grid), which is what scales past nx ~ 10^3.
    transform_fn(test_xs, X, Y) -> colouring factor A, optional (tests inject LAPACK's factor; default: device).
    Returns a dict with the arrays the reference saves as <criterion>_{xx,ff,yy,guess_xx,guess_ff}.npy plus the
    per-query test-set sizes and query indices (-1: the query was a lone maximizer, no criterion evaluated)."""
    from . import optfunc
    from .criteria import evaluate_mp, evaluate_sample_mp
    if criterion not in ("ftl", "sftl"):
        raise ValueError("only the TES criteria 'ftl' (TES_ep) and 'sftl' (TES_sp) are on this path")
    if nhyp != 1:
        raise ValueError("bo_discrete.py implements known GP hyper-parameters only (nhyp = 1, :550)")
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    nx, xdim = xs.shape
    xs_d = dev(xs)
    ls = np.asarray(l, dtype=np.float64).reshape(1, xdim)
    sigmas, sigma0s = np.array([float(sigma)]), np.array([float(sigma0)])
    st_mode = "sample" if criterion == "sftl" else mode
    out = dict(xx=np.zeros([nrun, nquery + ninit, xdim]), ff=np.zeros([nrun, nquery + ninit]),
               yy=np.zeros([nrun, nquery + ninit]), guess_xx=np.zeros([nrun, nquery + 1, xdim]),
               guess_ff=np.zeros([nrun, nquery + 1]), ntest=np.zeros([nrun, nquery], dtype=np.int64),
               query_idx=-np.ones([nrun, nquery], dtype=np.int64))

    def best_guess(X, Y, invK):
        mean, _ = ops.gp_mean_var(xs_d, dev(X), Y, ls[0], float(sigma), dev(invK)[0])
        return xs[int(ops.argmax(mean)[0])]

    for nr in range(nrun):
        np.random.seed(seed + nr)
        idx0 = np.random.randint(nx, size=ninit)
        X = xs[idx0, :]
        F = np.asarray(f(X), dtype=np.float64).reshape(-1, 1)
        Y = F + np.random.randn(X.shape[0], 1) * np.sqrt(sigma0)
        for nq in range(nquery):
            invK = utils.precomputeInvKs(xdim, 1, ls, sigmas, sigma0s, X)
            # ---- maximizer samples
            if sampler == "exact":
                z = np.random.randn(nmax, nx)
                max_xs = sample_maximizers(xs, X, Y, ls[0], sigma, sigma0, z)[0]
            else:
                theta, W, b = optfunc.draw_random_init_weights_features(xdim, nmax, n_features, X, Y, ls, sigma, sigma0)
                idx, _ = ops.rff_topk(xs_d, W, b, theta, float(sigma), k=1)
                max_xs = xs[idx[:, 0].cpu().numpy()]
            query_x = None
            uniq = utils.remove_duplicates_np(max_xs, resolution=duplicate_resolution)
            if uniq.shape[0] == 1:
                query_x = uniq
            else:
                test_xs = uniq.reshape(-1, xdim)
                if nmax > test_xs.shape[0]:                       # max_n_test = nmax (:880, :575-583)
                    n_extra = min(nmax - test_xs.shape[0], X.shape[0])
                    idxs = np.array(list(range(X.shape[0])))
                    np.random.shuffle(idxs)
                    test_xs = np.concatenate([test_xs, X[idxs[:n_extra], :]], axis=0)
                test_xs = utils.remove_duplicates_np(test_xs, resolution=duplicate_resolution)
                ntest = test_xs.shape[0]
                if ntest == 1:
                    query_x = test_xs
            if query_x is None:
                mean_t, cov_t = utils.compute_mean_var_f(test_xs, X, Y, ls[0], sigma, sigma0, invK[0], fullcov=True)
                tr = None if transform_fn is None else [transform_fn(test_xs, X, Y)]
                test_xs, max_probs, _, _, v0, v1 = utils.get_testidxs_stats(
                    1, test_xs, mean_t.reshape(1, ntest), cov_t.reshape(1, ntest, ntest), mode=st_mode,
                    nsample=ntestsample, transforms=tr)
                out["ntest"][nr, nq] = test_xs.shape[0]
                if test_xs.shape[0] == 1:
                    query_x = test_xs
            if query_x is None:
                _, invpNK = evaluate_mp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)
                cseed = (seed + nr) * 100003 + nq
                if criterion == "ftl":
                    acq = evaluate_mp.mp(xs_d, ls, sigmas, sigma0s, X, Y, xdim, nx, X.shape[0], 1, nysample, test_xs,
                                         max_probs, v0, v1, invK, invpNK, stochastic=1 - deterministic,
                                         niteration=nsto, seed=cseed)
                else:
                    acq = evaluate_sample_mp.mp(xs_d, ls, sigmas, sigma0s, X, Y, xdim, nx, X.shape[0], 1, nysample,
                                                test_xs, max_probs, v0, v1.astype(bool), invpNK, niteration=nsto,
                                                seed=cseed)
                qi = int(ops.argmax(acq)[0])                     # first maximum (bo_discrete.py:806)
                out["query_idx"][nr, nq] = qi
                query_x = xs[qi:qi + 1]
            # ---- best guess before the query is observed (:971-984), then observe (:1057-1075)
            g = best_guess(X, Y, invK)
            out["guess_xx"][nr, nq] = g
            out["guess_ff"][nr, nq] = np.asarray(f(g.reshape(-1, xdim))).squeeze()
            query_x = np.asarray(query_x, dtype=np.float64).reshape(-1, xdim)
            query_f = np.asarray(f(query_x), dtype=np.float64).reshape(-1, 1)
            query_y = query_f + np.random.randn(query_x.shape[0], 1) * np.sqrt(sigma0)
            if verbose:
                print("%d:%d QUERY %s  guess f=%.6g" % (nr, nq, query_x.squeeze(), out["guess_ff"][nr, nq]), flush=True)
            X = np.concatenate([X, query_x], axis=0)
            F = np.concatenate([F, query_f], axis=0)
            Y = np.concatenate([Y, query_y], axis=0)
        invK = utils.precomputeInvKs(xdim, 1, ls, sigmas, sigma0s, X)
        g = best_guess(X, Y, invK)
        out["guess_xx"][nr, nquery] = g
        out["guess_ff"][nr, nquery] = np.asarray(f(g.reshape(-1, xdim))).squeeze()
        out["xx"][nr], out["ff"][nr], out["yy"][nr] = X, F.squeeze(), Y.squeeze()
    return out


def save(out, folder, criterion):
    """The .npy files bo_discrete.py:1137-1143 writes (regret files excepted: not on the TES path)."""
    import os
    os.makedirs(folder, exist_ok=True)
    for key in ("xx", "ff", "yy", "guess_ff", "guess_xx"):
        np.save("{}/{}_{}.npy".format(folder, criterion, key), out[key])
