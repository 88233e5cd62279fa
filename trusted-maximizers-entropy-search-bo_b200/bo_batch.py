"""The continuous-domain batch BO loop of bo_batch.py:881-1125 for the TES criteria, re-hosted on the device kernels
(no TF session).  Per query, as `get_placeholder_values` (bo_batch.py:550-767) and the main loop (:975-1075) do:

  1. invKs; best guess = maximiser of the posterior mean (optimize_continuous_function on the candidate grid, :420-443)
  2. RFF posterior function samples (optfunc.draw_random_init_weights_features_np, :641-652, host RNG in the
     reference's order) and their maximisers (sample_xmaxs_fmaxs: top-k over the candidates + `ntrain` Adam steps)
  3. test_xs = the nmax maximisers (no de-duplication in this driver), posterior mean / covariance,
     utils.get_testidxs_stats(mode, nsample = 1000, n_min_sample = 2), invpNKs
  4. five random start batches drawn from test_xs (:1009-1043), the criterion's optimize_continuous_function with
     top_init_k = 3 and `ntrain` Adam steps (:1045-1058), query = its maximiser
  5. observe y = f(x) + N(0, sigma0) for the q points of the batch

Criteria (bo_batch.py `evaluate_criterion`): 'sftl' -> evaluate_batch_sample_mp.mp (batch TES_sp), 'ftl' ->
evaluate_batch_mp.mp (batch TES_ep, identically -(sum p) log(sum p): SURVEY Q5).  Known hyper-parameters only: the GPy
refit of :924-951 (>= 12 observations) is out of scope.

What cannot be replayed from numpy, because the reference draws it inside the TF graph: the top_init_k uniform
candidates appended by optimize_continuous_function (np.random.rand here) and the criterion's y draws (counter-based,
keyed on seed + run, query and Adam step; SHARED by all points scored in one step so the finite-difference gradient
uses common random numbers -- the reference back-propagates through one reparameterised draw per step, the same
estimator up to O(h^2)).
"""
import numpy as np

from . import ops, optfunc, utils, utils_for_continuous
from .device import dev

perturb_std = 1e-2  # bo_batch.py:26


def evaluate_batch_mp_mod():
    """The criteria module whose get_pNK_test_obs the loop calls (a hook point for the tests)."""
    from .criteria import evaluate_batch_mp
    return evaluate_batch_mp


def crit_start_batches(pool, batchsize, xmin, xmax, n_crit_candidate_xs=5):
    """bo_batch.py:1009-1043: `n_crit_candidate_xs` start batches of `batchsize` points drawn from `pool`
    (np.random.shuffle; a pool smaller than the batch is recycled and the repeats perturbed)."""
    xdim = pool.shape[1]
    out = np.zeros([n_crit_candidate_xs, batchsize, xdim])
    idxs = np.array(list(range(pool.shape[0]))).astype(int)
    for ci in range(n_crit_candidate_xs):
        np.random.shuffle(idxs)
        selected = idxs[:batchsize]
        is_duplicated = False
        while selected.shape[0] < batchsize:
            is_duplicated = True
            selected = np.concatenate([selected, idxs[:(batchsize - selected.shape[0])]], axis=0)
        an_init = pool[selected, :].reshape(batchsize, xdim)
        if is_duplicated:
            an_init[idxs.shape[0]:] += np.clip(np.random.randn(batchsize - idxs.shape[0], xdim) * perturb_std,
                                               a_min=-perturb_std * 2, a_max=perturb_std * 2)
            an_init = np.clip(an_init, a_min=xmin, a_max=xmax)
        out[ci, ...] = an_init
    return out


def run(f, xs, xmin, xmax, l, sigma, sigma0, criterion="sftl", mode="sample", nquery=3, nrun=1, ninit=2, nmax=30,
        nfeature=300, nsto=10, ntrain=300, nysample=100, batchsize=1, ntestsample=1000, n_min_sample=2, seed=1,
        top_init_k=3, fd_step=1e-5, transform_fn=None, verbose=False, trace=None, single=False, deterministic=0,
        duplicate_resolution=1e-3, max_nan_retry=10):
    """f: objective on (n, xdim) arrays; xs: candidate grid (`f_info['xs']`, bo_batch.py:783) used to start the
    optimisation of the posterior mean and of the function samples; flags as in bo_batch.py (`--criterion --mode
    --numqueries --numruns --ninit --nmax --nfeature --nsto --ntrain --nysample --batchsize`).
    transform_fn(test_xs, X, Y) -> colouring factor for get_testidxs_stats (tests inject LAPACK's; default: device).
    single=True: the sequential driver bo.py instead (bo.py:1026-1244; see tes_b200/bo.py): batchsize 1, maximisers
    de-duplicated at `duplicate_resolution` (bo.py:830-836), criteria evaluate_sample_mp.mp ('sftl') / evaluate_mp.mp
    ('ftl', stochastic = 1 - deterministic), criterion start points = all nmax maximisers (bo.py:1151-1153), lone
    maximiser queried directly, and the whole query repeated when it comes out NaN (bo.py:1190-1193).
    `trace`: optional list that receives one dict of intermediates per query (maximizers, test points, start batches,
    their criterion values, the refined query).
    Returns the arrays the reference saves (xx, ff, yy, guess_xx, guess_ff) plus per-query diagnostics."""
    from .criteria import evaluate_batch_sample_mp, evaluate_mp, evaluate_sample_mp
    if single and int(batchsize) != 1:
        raise ValueError("bo.py queries one point at a time")
    if criterion not in ("ftl", "sftl"):
        raise ValueError("only the TES criteria 'ftl' and 'sftl' are on this path")
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    xdim = xs.shape[1]
    q = int(batchsize)
    ls = np.asarray(l, dtype=np.float64).reshape(1, xdim)
    sigmas, sigma0s = np.array([float(sigma)]), np.array([float(sigma0)])
    st_mode = "sample" if criterion == "sftl" else mode
    ntot = nquery * q + ninit
    out = dict(xx=np.zeros([nrun, ntot, xdim]), ff=np.zeros([nrun, ntot]), yy=np.zeros([nrun, ntot]),
               guess_xx=np.zeros([nrun, nquery + 1, xdim]), guess_ff=np.zeros([nrun, nquery + 1]),
               ntest=np.zeros([nrun, nquery], dtype=np.int64), crit_max=np.full([nrun, nquery], np.nan))

    def best_guess(X, Y, invK):
        Xd, invd = dev(X), dev(invK)
        func = lambda x, step: utils.compute_mean_f(x, xdim, 1, Xd, Y, ls, sigmas, sigma0s, invd)
        return utils_for_continuous.optimize_continuous_function(xdim, func, xs, top_init_k, xmin, xmax, ntrain=ntrain,
                                                                 fd_step=fd_step)[0]

    for nr in range(nrun):
        np.random.seed(seed + nr)
        X = np.random.rand(ninit, xdim) * (xmax - xmin) + xmin
        F = np.asarray(f(X), dtype=np.float64).reshape(-1, 1)
        Y = F + np.random.randn(X.shape[0], 1) * np.sqrt(sigma0)
        for nq in range(nquery):
            for _retry in range(max_nan_retry + 1):      # bo.py:1190-1193 / bo_batch.py:1080-1083: repeat on NaN
                invK = utils.precomputeInvKs(xdim, 1, ls, sigmas, sigma0s, X)
                g = best_guess(X, Y, invK)
                out["guess_xx"][nr, nq] = g.squeeze()
                out["guess_ff"][nr, nq] = np.asarray(f(g.reshape(1, xdim))).squeeze()
                # ---- function samples and their maximisers
                theta, W, b = optfunc.draw_random_init_weights_features_np(xdim, nmax, nfeature, X, Y, ls, sigma, sigma0)
                lo = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (xdim,))
                hi = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (xdim,))
                cand = np.concatenate([xs, lo + np.random.rand(top_init_k, xdim) * (hi - lo)], axis=0)
                _, max_xs, _, _ = utils_for_continuous.sample_xmaxs_fmaxs(
                    xdim, 1, nmax, nfeature, ls, sigmas, sigma0s, theta[None], W[None], b[None], cand, top_init_k,
                    xmin=xmin, xmax=xmax, ntrain=ntrain)
                maximizers = max_xs[0].reshape(-1, xdim)
                tr_rec = dict(guess=g.copy(), maximizers=maximizers.copy())
                if trace is not None:
                    trace.append(tr_rec)
                test_xs, query_x = maximizers, None
                if single:
                    test_xs = utils.remove_duplicates_np(maximizers, resolution=duplicate_resolution).reshape(-1, xdim)
                lone = (lambda: test_xs) if single else (lambda: maximizers[:q])   # bo.py:843-845 / bo_batch.py:710
                ntest = test_xs.shape[0]
                if ntest == 1:
                    query_x = lone()
                else:
                    mean_t, cov_t = utils.compute_mean_var_f(test_xs, X, Y, ls[0], sigma, sigma0, invK[0], fullcov=True)
                    tr = None if transform_fn is None else [transform_fn(test_xs, X, Y)]
                    test_xs, max_probs, _, _, v0, v1 = utils.get_testidxs_stats(
                        1, test_xs, mean_t.reshape(1, ntest), cov_t.reshape(1, ntest, ntest), mode=st_mode,
                        nsample=ntestsample, n_min_sample=n_min_sample, transforms=tr)
                    out["ntest"][nr, nq] = test_xs.shape[0]
                    if test_xs.shape[0] == 1:
                        query_x = lone()
                if query_x is None:
                    _, invpNK = evaluate_batch_mp_mod().get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)
                    starts = maximizers.reshape(-1, 1, xdim) if single else crit_start_batches(test_xs, q, xmin, xmax)
                    cseed = ((seed + nr) * 100003 + nq) * 100003
                    Xd, Yd, tx, mp_, v0d, inv_d = dev(X), dev(Y), dev(test_xs), dev(max_probs), dev(v0), dev(invpNK)
                    v1b = np.asarray(v1).astype(bool) if criterion == "sftl" else None

                    def crit(xflat, step):
                        xb = np.asarray(xflat, dtype=np.float64).reshape(-1, q, xdim)
                        if single and criterion == "sftl":
                            return evaluate_sample_mp.mp(dev(xb[:, 0]), ls, sigmas, sigma0s, Xd, Yd, xdim, xb.shape[0],
                                                         X.shape[0], 1, nysample, tx, mp_, v0d, v1b, inv_d, niteration=nsto,
                                                         seed=cseed + step, n_global=-1).cpu().numpy()
                        if single:
                            return evaluate_mp.mp(dev(xb[:, 0]), ls, sigmas, sigma0s, Xd, Yd, xdim, xb.shape[0], X.shape[0], 1,
                                                  nysample, tx, mp_, v0d, dev(v1), dev(invK), inv_d,
                                                  stochastic=1 - deterministic, niteration=nsto, seed=cseed + step,
                                                  n_global=-1).cpu().numpy()
                        if criterion == "sftl":
                            return evaluate_batch_sample_mp.mp(dev(xb), ls, sigmas, sigma0s, Xd, Yd, xdim, X.shape[0], 1,
                                                               nysample, tx, mp_, v0d, v1b, inv_d, niteration=nsto,
                                                               seed=cseed + step, n_global=-1).cpu().numpy()
                        return ops.batch_ep_acq(mp_[0][mp_[0] > 0], dev(xb)).cpu().numpy()
                    qx, qv = utils_for_continuous.optimize_continuous_function(
                        xdim * q, crit, starts.reshape(-1, xdim * q), top_init_k, np.tile(lo, q), np.tile(hi, q),
                        ntrain=ntrain, fd_step=fd_step)
                    out["crit_max"][nr, nq] = qv
                    query_x = qx.reshape(q, xdim)
                    tr_rec.update(test_xs=np.asarray(test_xs).copy(), max_probs=np.asarray(max_probs).copy(), starts=starts,
                                  start_vals=crit(starts.reshape(-1, xdim * q), 0), qx=qx, qv=qv)
                if query_x is not None and not np.any(np.isnan(np.asarray(query_x, dtype=np.float64))):
                    break
            else:
                # the reference repeats until the query is finite (bo_batch.py:1080-1081); a NaN query must never be
                # observed -- it would poison every later GP solve
                raise ops._lib.TesError("query %d of run %d is still NaN after %d attempts" % (nq, nr, max_nan_retry + 1))
            query_x = np.asarray(query_x, dtype=np.float64).reshape(q, xdim)
            query_f = np.asarray(f(query_x), dtype=np.float64).reshape(q, 1)
            query_y = query_f + np.random.randn(q, 1) * np.sqrt(sigma0)
            if verbose:
                print("%d:%d QUERY %s  guess f=%.6g" % (nr, nq, query_x.reshape(-1), out["guess_ff"][nr, nq]), flush=True)
            X = np.concatenate([X, query_x], axis=0)
            F = np.concatenate([F, query_f], axis=0)
            Y = np.concatenate([Y, query_y], axis=0)
        invK = utils.precomputeInvKs(xdim, 1, ls, sigmas, sigma0s, X)
        g = best_guess(X, Y, invK)
        out["guess_xx"][nr, nquery] = g.squeeze()
        out["guess_ff"][nr, nquery] = np.asarray(f(g.reshape(1, xdim))).squeeze()
        out["xx"][nr], out["ff"][nr], out["yy"][nr] = X, F.squeeze(), Y.squeeze()
    return out


def save(out, folder, criterion):
    """The .npy files bo_batch.py:1113-1119 writes."""
    import os
    os.makedirs(folder, exist_ok=True)
    for key in ("xx", "ff", "yy", "guess_ff", "guess_xx"):
        np.save("{}/{}_{}.npy".format(folder, criterion, key), out[key])
