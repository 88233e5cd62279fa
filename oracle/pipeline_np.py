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
             duplicate_resolution=1e-3, t_max=None, return_acq=False, transform=None):
    N, d = cand.shape
    n = X.shape[0]
    Y = Y.reshape(-1, 1)
    ls, sigmas, sigma0s = np.asarray(l, dtype=np.float64).reshape(1, d), np.array([sigma]), np.array([sigma0])
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
    if criterion == "ftl":
        acq = onp.evaluate_mp(cand, ls, sigmas, sigma0s, X, Y, d, N, n, 1, nysample, test_k, max_probs, v0, v1,
                              invK, invpNK, stochastic=not deterministic, niteration=niteration, z=z)
    else:
        out["K"] = int(v0.shape[-1])
        acq = onp.evaluate_sample_mp(cand, ls, sigmas, sigma0s, X, Y, d, N, n, 1, nysample, test_k, max_probs, v0,
                                     v1, invpNK, niteration=niteration, z=z)
    qidx = int(np.argmax(acq))
    out.update(query_idx=qidx, query_val=float(acq[qidx]), query_x=cand[qidx])
    if return_acq:
        out["acq"] = acq
    return out
