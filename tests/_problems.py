"""Seeded small TES problems shared by the parity tests (built with the numpy oracle only)."""
import numpy as np

from oracle import tes_oracle_np as onp

BRANIN = dict(l=np.array([25.5866, 6.3735]), sigma=0.589, sigma0=1e-4)          # functions.py:571-577
PH = dict(l=np.array([177.9418631280815, 309.4630784148164]), sigma=0.29444226420446995,
          sigma0=0.06476473751595126)                                            # pHdata/hyperparameters_pH.txt
HART6 = dict(l=np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]), sigma=1.423, sigma0=1e-4)


def make_problem(seed=0, d=2, n=5, T=6, nx=7, hyp=BRANIN, nsample=300, mode="sample", n_min_sample=2,
                 sigma0=None):
    """Returns a dict with everything a criterion call needs (nhyp = 1)."""
    rs = np.random.RandomState(seed)
    l, sigma = hyp["l"][:d] if len(hyp["l"]) >= d else np.resize(hyp["l"], d), hyp["sigma"]
    s0 = hyp["sigma0"] if sigma0 is None else sigma0
    X = rs.rand(n, d)
    Y = rs.randn(n, 1) * 0.5
    test_xs = rs.rand(T, d)
    x = rs.rand(nx, d)
    ls, sigmas, sigma0s = l.reshape(1, d), np.array([sigma]), np.array([s0])
    invK = onp.precomputeInvKs(ls, sigmas, sigma0s, X)
    mean, cov = onp.compute_mean_var_f(test_xs, X, Y, l, sigma, s0, invK[0], fullcov=True)
    z = rs.normal(size=(1, nsample, T, 1))
    out = onp.get_testidxs_stats(1, test_xs.copy(), mean.reshape(1, T).copy(), cov.reshape(1, T, T).copy(),
                                 mode=mode, nsample=nsample, n_min_sample=n_min_sample, z=z)
    test_k, max_probs, mean_k, cov_k, v0, v1 = out
    _, invpNK = onp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_k)
    return dict(x=x, X=X, Y=Y, d=d, n=n, ls=ls, sigmas=sigmas, sigma0s=sigma0s, l=l, sigma=sigma, sigma0=s0,
                test_xs=test_k, max_probs=max_probs, test_mean=mean_k, test_cov=cov_k, v0=v0, v1=v1,
                invpNK=invpNK, invK=invK, T=test_k.shape[0], nx=nx, rs=rs)
