"""TEST INFRASTRUCTURE ONLY -- CPU oracle of one full TES acquisition step, composed from
oracle/tes_oracle_np.py and oracle/tes_oracle.c in the order the reference's drivers run it
(bo_discrete.py:504-647 + :1032-1055; bo_batch.py:550-767 + :994-1061).  Used by the end-to-end
parity tests and as the CPU baseline of bench.py."""
import numpy as np

from . import c_oracle as co
from . import tes_oracle_np as onp


def select_test_points(max_xs, resolution=1e-3, t_max=None):
    uniq = onp.remove_duplicates(max_xs, resolution)
    if t_max is not None and uniq.shape[0] > t_max:
        d = np.sqrt(((max_xs[:, None, :] - uniq[None, :, :]) ** 2).sum(-1))
        owner = np.argmax(d <= resolution, axis=1)
        counts = np.bincount(owner, minlength=uniq.shape[0])
        order = np.lexsort((np.arange(uniq.shape[0]), -counts))[:t_max]
        uniq = uniq[np.sort(order)]
    return uniq


def tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, criterion="ftl", mode="empirical", deterministic=True,
             ntestsample=1000, n_min_sample=2, nysample=10, niteration=10, z_joint=None, z=None, seed=0,
             duplicate_resolution=1e-3, t_max=None, return_acq=False, transform=None, q=1, draws=None):
    """`draws` = (randn_W, rand_b, randn_noise): the step starts at the RFF posterior weight draw
    (optfunc.py:303-385, host numpy in the reference) and theta/W/b are ignored.
    `z` None with a Monte-Carlo criterion: the counter-based draws of include/tes_spec.h (the streams the
    kernels key on the global candidate index), so both paths consume identical normals.  `q` > 1: batch TES_sp
    on every q consecutive candidates (evaluate_batch_sample_mp.py)."""
    N, d = cand.shape
    n = X.shape[0]
    Y = Y.reshape(-1, 1)
    ls, sigmas, sigma0s = np.asarray(l, dtype=np.float64).reshape(1, d), np.array([sigma]), np.array([sigma0])
    if draws is not None:
        S_, F_ = draws[0].shape[0], draws[0].shape[1]
        theta, W, b = onp.draw_random_init_weights_features(d, S_, F_, X, Y, ls, sigma, sigma0, *draws)
    idx, val = co.rff_topk(cand, W, b, theta, sigma, 1)
    test_xs = select_test_points(cand[idx[:, 0]], duplicate_resolution, t_max)
    out = {"n_unique_maximizers": int(test_xs.shape[0]), "argmax_idx": idx[:, 0], "argmax_val": val[:, 0]}
    if test_xs.shape[0] == 1:
        out.update(query_x=test_xs[0], query_idx=int(idx[0, 0]), query_val=float("nan"), T=1, Sp=1)
        return out
    invK = onp.precomputeInvKs(ls, sigmas, sigma0s, X)
    mean_t, cov_t = onp.compute_mean_var_f(test_xs, X, Y, ls[0], sigma, sigma0, invK[0], fullcov=True)
    T0 = test_xs.shape[0]
    if z_joint is None:
        # include/tes_spec.h: counter-based draws, stream TES_STREAM_JOINT_Z = 5, iteration = hyper-parameter index
        z_joint = co.normals(seed, 0, 5, 0, ntestsample * T0).reshape(1, ntestsample, T0, 1)
    st_mode = "sample" if criterion == "sftl" else mode
    test_k, max_probs, mean_k, cov_k, v0, v1 = onp.get_testidxs_stats(
        1, test_xs, mean_t.reshape(1, T0), cov_t.reshape(1, T0, T0), mode=st_mode, nsample=ntestsample,
        n_min_sample=n_min_sample, z=z_joint, transforms=None if transform is None else [transform])
    out.update(T=int(test_k.shape[0]), Sp=int((max_probs[0] > 0).sum()), test_cov=cov_t)
    if test_k.shape[0] == 1:
        out.update(query_x=test_k[0], query_idx=-1, query_val=float("nan"))
        return out
    _, invpNK = onp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_k)
    Sp = out["Sp"]

    def philox(stream, shape):  # z[it][iy][j][x]... of tes_spec.h, one iteration per Philox `iteration` counter
        per_it = int(np.prod(shape))
        return np.stack([co.normals(seed, it, stream, 0, per_it).reshape(shape) for it in range(niteration)])[None]
    if criterion == "ftl":
        if z is None and not deterministic:
            z = philox(1, (nysample, Sp, N))
        acq = onp.evaluate_mp(cand, ls, sigmas, sigma0s, X, Y, d, N, n, 1, nysample, test_k, max_probs, v0, v1,
                              invK, invpNK, stochastic=not deterministic, niteration=niteration, z=z)
    elif q > 1:
        out["K"] = K = int(v0.shape[-1])
        nx = N // q
        if z is None:
            z = philox(3, (nysample, Sp, nx, K, q))
        acq = onp.evaluate_batch_sample_mp(cand.reshape(nx, q, d), ls, sigmas, sigma0s, X, Y, d, n, 1, nysample, test_k,
                                           max_probs, v0, v1, invpNK, niteration=niteration, z=z)
    else:
        out["K"] = K = int(v0.shape[-1])
        if z is None:
            z = philox(2, (nysample, Sp, N, K))
        acq = onp.evaluate_sample_mp(cand, ls, sigmas, sigma0s, X, Y, d, N, n, 1, nysample, test_k, max_probs, v0,
                                     v1, invpNK, niteration=niteration, z=z)
    qidx = int(np.argmax(acq))
    out.update(query_idx=qidx, query_val=float(acq[qidx]),
               query_x=cand[qidx] if q == 1 else cand[qidx * q:(qidx + 1) * q])
    if return_acq:
        out["acq"] = acq
    return out


def run_discrete_bo(f, xs, l, sigma, sigma0, criterion="ftl", mode="ep", nquery=50, nrun=1, ninit=2, nmax=5,
                    deterministic=0, nsto=50, nysample=10, ntestsample=1000, seed=1, duplicate_resolution=1e-3,
                    transform_fn=None):
    """TEST INFRASTRUCTURE -- numpy restatement of the BO loop of bo_discrete.py:928-1145 for the TES criteria with
    the exact discrete maximizer sampler (:359-400), consuming the host RNG in the same order as the device loop
    (tes_b200.bo_discrete.run) and the same counter-based draws for the stochastic criteria.  Returns the same dict."""
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    nx, xdim = xs.shape
    ls, sigmas, sigma0s = np.asarray(l, dtype=np.float64).reshape(1, xdim), np.array([float(sigma)]), np.array([float(sigma0)])
    st_mode = "sample" if criterion == "sftl" else mode
    out = dict(xx=np.zeros([nrun, nquery + ninit, xdim]), ff=np.zeros([nrun, nquery + ninit]),
               yy=np.zeros([nrun, nquery + ninit]), guess_xx=np.zeros([nrun, nquery + 1, xdim]),
               guess_ff=np.zeros([nrun, nquery + 1]), ntest=np.zeros([nrun, nquery], dtype=np.int64),
               query_idx=-np.ones([nrun, nquery], dtype=np.int64))

    def best_guess(X, Y, invK):
        mean, _ = onp.compute_mean_var_f(xs, X, Y, ls[0], sigma, sigma0, invK[0], fullcov=False)
        return xs[int(np.argmax(mean))]

    for nr in range(nrun):
        np.random.seed(seed + nr)
        idx0 = np.random.randint(nx, size=ninit)
        X = xs[idx0, :]
        F = np.asarray(f(X), dtype=np.float64).reshape(-1, 1)
        Y = F + np.random.randn(X.shape[0], 1) * np.sqrt(sigma0)
        for nq in range(nquery):
            invK = onp.precomputeInvKs(ls, sigmas, sigma0s, X)
            z = np.random.randn(nmax, nx)
            midx, _ = onp.discrete_maximizer_samples(xs, X, Y, ls[0], sigma, sigma0, z)
            max_xs = xs[midx]
            query_x = None
            uniq = onp.remove_duplicates(max_xs, duplicate_resolution)
            if uniq.shape[0] == 1:
                query_x = uniq
            else:
                test_xs = uniq.reshape(-1, xdim)
                if nmax > test_xs.shape[0]:
                    n_extra = min(nmax - test_xs.shape[0], X.shape[0])
                    idxs = np.array(list(range(X.shape[0])))
                    np.random.shuffle(idxs)
                    test_xs = np.concatenate([test_xs, X[idxs[:n_extra], :]], axis=0)
                test_xs = onp.remove_duplicates(test_xs, duplicate_resolution)
                ntest = test_xs.shape[0]
                if ntest == 1:
                    query_x = test_xs
            if query_x is None:
                mean_t, cov_t = onp.compute_mean_var_f(test_xs, X, Y, ls[0], sigma, sigma0, invK[0], fullcov=True)
                tr = None if transform_fn is None else [transform_fn(test_xs, X, Y)]
                zj = np.random.normal(size=(1, ntestsample, ntest, 1))          # utils.py:1674
                test_xs, max_probs, _, _, v0, v1 = onp.get_testidxs_stats(
                    1, test_xs, mean_t.reshape(1, ntest), cov_t.reshape(1, ntest, ntest), mode=st_mode,
                    nsample=ntestsample, z=zj, transforms=tr)
                out["ntest"][nr, nq] = test_xs.shape[0]
                if test_xs.shape[0] == 1:
                    query_x = test_xs
            if query_x is None:
                _, invpNK = onp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)
                cseed = (seed + nr) * 100003 + nq
                Sp, n = int((max_probs[0] > 0).sum()), X.shape[0]
                if criterion == "ftl":
                    zc = None
                    if not deterministic:
                        zc = np.stack([co.normals(cseed, it, 1, 0, nysample * Sp * nx).reshape(nysample, Sp, nx)
                                       for it in range(nsto)])[None]
                    acq = onp.evaluate_mp(xs, ls, sigmas, sigma0s, X, Y, xdim, nx, n, 1, nysample, test_xs, max_probs,
                                          v0, v1, invK, invpNK, stochastic=not deterministic, niteration=nsto, z=zc)
                else:
                    K = v0.shape[-1]
                    zc = np.stack([co.normals(cseed, it, 2, 0, nysample * Sp * nx * K).reshape(nysample, Sp, nx, K)
                                   for it in range(nsto)])[None]
                    acq = onp.evaluate_sample_mp(xs, ls, sigmas, sigma0s, X, Y, xdim, nx, n, 1, nysample, test_xs,
                                                 max_probs, v0, v1, invpNK, niteration=nsto, z=zc)
                qi = int(np.argmax(acq))
                out["query_idx"][nr, nq] = qi
                query_x = xs[qi:qi + 1]
            g = best_guess(X, Y, invK)
            out["guess_xx"][nr, nq] = g
            out["guess_ff"][nr, nq] = np.asarray(f(g.reshape(-1, xdim))).squeeze()
            query_x = np.asarray(query_x, dtype=np.float64).reshape(-1, xdim)
            query_f = np.asarray(f(query_x), dtype=np.float64).reshape(-1, 1)
            query_y = query_f + np.random.randn(query_x.shape[0], 1) * np.sqrt(sigma0)
            X = np.concatenate([X, query_x], axis=0)
            F = np.concatenate([F, query_f], axis=0)
            Y = np.concatenate([Y, query_y], axis=0)
        invK = onp.precomputeInvKs(ls, sigmas, sigma0s, X)
        g = best_guess(X, Y, invK)
        out["guess_xx"][nr, nquery] = g
        out["guess_ff"][nr, nquery] = np.asarray(f(g.reshape(-1, xdim))).squeeze()
        out["xx"][nr], out["ff"][nr], out["yy"][nr] = X, F.squeeze(), Y.squeeze()
    return out
