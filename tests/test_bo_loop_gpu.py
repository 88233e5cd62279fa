"""The re-hosted BO loop (tes_b200.bo_discrete.run; reference: bo_discrete.py:928-1145) against its numpy
restatement over several queries on the Brooms-Barn pH grid (BASELINE config C2: shipped hyper-parameters, 20x20
grid, nmax = 5, TES_ep with modes ep / empirical, and TES_sp): the whole trajectory -- initial design, every query
point, every best guess -- must be IDENTICAL, with the host RNG consumed in the same order on both sides."""
import numpy as np
import pytest

from oracle import pipeline_np
from oracle import tes_oracle_np as onp

pytestmark = pytest.mark.gpu


def _problem(golden):
    g = golden("kat_objectives")
    hyp = g["pH_hyp"]
    xs, fx = g["pH_xs"], g["pH_f_on_xs"]

    def f(x):  # the pH objective restricted to its grid (functions.py:367-398): nearest grid point's value
        x = np.asarray(x).reshape(-1, xs.shape[1])
        return fx[np.argmin(((x[:, None, :] - xs[None, :, :]) ** 2).sum(-1), axis=1)]
    return f, xs, hyp[1:3], float(hyp[0]), float(hyp[3])


def _lapack_factor(l, sigma, sigma0):
    def fn(test_xs, X, Y):
        invK = onp.precomputeInvKs(np.asarray(l).reshape(1, -1), np.array([sigma]), np.array([sigma0]), X)
        _, cov = onp.compute_mean_var_f(test_xs, X, Y, l, sigma, sigma0, invK[0], fullcov=True)
        return onp.mvn_transform_matrix(cov)
    return fn


@pytest.mark.parametrize("criterion,mode,det,kw", [
    ("ftl", "ep", 1, dict(nquery=6)),
    ("ftl", "empirical", 1, dict(nquery=6)),
    ("ftl", "empirical", 0, dict(nquery=4, nsto=3, nysample=4)),
    ("sftl", "sample", 0, dict(nquery=3, nsto=2, nysample=3, ntestsample=200)),
])
def test_discrete_bo_trajectory_is_identical(golden, criterion, mode, det, kw):
    from tes_b200 import bo_discrete
    f, xs, l, sigma, sigma0 = _problem(golden)
    args = dict(criterion=criterion, mode=mode, deterministic=det, nrun=1, ninit=2, nmax=5, seed=1,
                transform_fn=_lapack_factor(l, sigma, sigma0), **kw)
    ref = pipeline_np.run_discrete_bo(f, xs, l, sigma, sigma0, **args)
    out = bo_discrete.run(f, xs, l, sigma, sigma0, **args)
    np.testing.assert_array_equal(out["query_idx"], ref["query_idx"])
    np.testing.assert_array_equal(out["ntest"], ref["ntest"])
    for key in ("xx", "ff", "yy", "guess_xx", "guess_ff"):
        np.testing.assert_array_equal(out[key], ref[key], err_msg=key)
    assert (ref["query_idx"] >= 0).sum() >= 2      # the criterion really decided some of the queries


def test_discrete_bo_runs_with_the_device_factor_and_rff_sampler(golden, tmp_path):
    """No injected factor (device eigen-solver) and the RFF maximizer sampler: a different realisation than the
    reference's LAPACK path, so only sanity is checked: shapes, observations on the grid, files written."""
    from tes_b200 import bo_discrete
    f, xs, l, sigma, sigma0 = _problem(golden)
    out = bo_discrete.run(f, xs, l, sigma, sigma0, criterion="ftl", mode="ep", deterministic=1, nquery=4, nrun=2,
                          ninit=2, nmax=5, seed=3, sampler="rff", n_features=300)
    assert out["xx"].shape == (2, 6, 2) and np.all(np.isfinite(out["yy"]))
    d = ((out["xx"].reshape(-1, 1, 2) - xs[None]) ** 2).sum(-1).min(1)
    assert np.all(d == 0.0)
    bo_discrete.save(out, str(tmp_path), "ftl")
    assert np.array_equal(np.load(str(tmp_path / "ftl_xx.npy")), out["xx"])


# ---- bo_batch.py loop (continuous domain, batch queries, Adam refinement of maximisers and criterion) ------------
class _OracleBackend:
    """Oracle-backed stand-ins with the signatures the re-hosted bo_batch loop calls, so the SAME control flow can
    be run on the numpy oracle and its trajectory compared with the device run."""

    def __init__(self):
        from oracle import c_oracle as co
        self.co = co

    # utils
    def precomputeInvKs(self, xdim, nhyp, ls, sigmas, sigma0s, X):
        return onp.precomputeInvKs(ls, sigmas, sigma0s, X)

    def compute_mean_f(self, x, xdim, n_hyp, X, Y, ls, sigmas, sigma0s, invKs):
        return onp.compute_mean_var_f(np.asarray(x), X, Y, ls[0], sigmas[0], sigma0s[0], invKs[0])[0].reshape(-1)

    def compute_mean_var_f(self, x, X, Y, l, sigma, sigma0, NKInv=None, fullcov=False):
        return onp.compute_mean_var_f(x, X, Y, l, sigma, sigma0, NKInv, fullcov=fullcov)

    def get_testidxs_stats(self, nhyp, test_xs, means, covs, mode="sample", nsample=1000, n_min_sample=2,
                           transforms=None):
        z = np.random.normal(size=(nsample, test_xs.shape[0], 1))[None]          # utils.py:1674
        return onp.get_testidxs_stats(nhyp, test_xs, means, covs, mode=mode, nsample=nsample, n_min_sample=n_min_sample,
                                      z=z, transforms=transforms)

    # optfunc
    def draw_random_init_weights_features_np(self, xdim, S, F, X, Y, l, sigma, sigma0):
        rW, rb, rn = np.random.randn(S, F, xdim), np.random.rand(S, F, 1), np.random.randn(S, F, 1)
        return onp.draw_random_init_weights_features(xdim, S, F, X, Y, l, sigma, sigma0, rW, rb, rn)

    # utils_for_continuous
    def sample_xmaxs_fmaxs(self, xdim, n_hyp, n_funcs, n_features, ls, sigmas, sigma0s, thetas, Ws, bs, init, k,
                           xmin=None, xmax=None, ntrain=0):
        idx, _ = self.co.rff_topk(init, Ws[0], bs[0], thetas[0], float(sigmas[0]), k)
        x0 = init[idx]
        xr, fr = onp.adam_refine_maximizers(x0, thetas[0], Ws[0], bs[0], float(sigmas[0]), xmin, xmax, ntrain)
        best = np.argmax(fr, axis=1)
        rows = np.arange(fr.shape[0])
        return x0[None], xr[rows, best][None], fr[rows, best][None], idx[None]

    # criteria.evaluate_batch_sample_mp.mp with SHARED counter-based draws (n_global = -1)
    def batch_sample_mp(self, x, ls, sigmas, sigma0s, X, Y, xdim, nobs, nhyp, nysample, test_xs, max_probs, v0, v1,
                        invpNK, niteration=10, seed=0, n_global=None):
        assert n_global == -1
        x = np.asarray(x)
        nx, q, _ = x.shape
        Sp, K = int((np.asarray(max_probs)[0] > 0).sum()), np.asarray(v0).shape[-1]
        z = np.stack([self.co.normals(seed, it, 2 if q == 1 else 3, 0, nysample * Sp * K * q)   # tes_spec.h streams
                     .reshape(nysample, Sp, 1, K, q)
                      for it in range(niteration)])
        z = np.broadcast_to(z, (niteration, nysample, Sp, nx, K, q))[None]
        return _Np(onp.evaluate_batch_sample_mp(x, ls, sigmas, sigma0s, np.asarray(X), np.asarray(Y), xdim, nobs, 1,
                                                nysample, np.asarray(test_xs), np.asarray(max_probs), np.asarray(v0),
                                                np.asarray(v1), np.asarray(invpNK), niteration=niteration, z=z))


    def _shared_z(self, seed, stream, niteration, shape_tail, nx):
        """Counter-based draws SHARED by all candidates (n_global = -1): z[it][iy][j][x]... with the x axis broadcast."""
        n = int(np.prod(shape_tail))
        return np.stack([self.co.normals(seed, it, stream, 0, n).reshape(shape_tail) for it in range(niteration)])

    # criteria.evaluate_sample_mp.mp (TES_sp, q = 1)
    def sample_mp(self, x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs, max_probs, v0, v1,
                  invpNK, niteration=10, seed=0, n_global=None):
        assert n_global == -1
        Sp, K = int((np.asarray(max_probs)[0] > 0).sum()), np.asarray(v0).shape[-1]
        z = self._shared_z(seed, 2, niteration, (nysample, Sp, 1, K), nx)
        z = np.broadcast_to(z, (niteration, nysample, Sp, nx, K))[None]
        return _Np(onp.evaluate_sample_mp(np.asarray(x), ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, 1, nysample, test_xs,
                                          max_probs, v0, v1, invpNK, niteration=niteration, z=z))

    # criteria.evaluate_mp.mp (TES_ep, q = 1)
    def ep_mp(self, x, ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, nhyp, nysample, test_xs, max_probs, pm, pc, invK,
              invpNK, stochastic=False, niteration=10, seed=0, n_global=None):
        assert n_global == -1
        Sp = int((np.asarray(max_probs)[0] > 0).sum())
        z = None
        if stochastic:
            z = self._shared_z(seed, 1, niteration, (nysample, Sp, 1), nx)
            z = np.broadcast_to(z, (niteration, nysample, Sp, nx))[None]
        return _Np(onp.evaluate_mp(np.asarray(x), ls, sigmas, sigma0s, X, Y, xdim, nx, nobs, 1, nysample, test_xs,
                                   max_probs, pm, pc, invK, invpNK, stochastic=bool(stochastic), niteration=niteration, z=z))

    def remove_duplicates_np(self, xs, resolution=1e-5):
        return onp.remove_duplicates(np.asarray(xs), resolution)


class _Np:  # .cpu().numpy() on a numpy result
    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


def _branin(golden):
    g = golden("kat_objectives")
    xs = g["branin_xs"]

    def f(x):  # functions.negative_Branin (functions.py:548-581), restated
        x = np.asarray(x).reshape(-1, 2)
        x1, x2 = 15.0 * x[:, 0] - 5.0, 15.0 * x[:, 1]
        v = (x2 - 5.1 / (4 * np.pi ** 2) * x1 ** 2 + 5.0 / np.pi * x1 - 6.0) ** 2 + 10.0 * (1 - 1 / (8 * np.pi)) * np.cos(x1) + 10.0
        return -(v - 54.8104) / 51.9496
    return f, xs, g["branin_l"], float(g["branin_sigma"]), float(g["branin_sigma0"])


def _stable_factor(l, sigma, sigma0):
    """A colouring factor that is a continuous function of its inputs (symmetric eigen-decomposition, ascending order,
    sign-normalised vectors): the two runs reach the test points through Adam paths that differ by rounding, and
    LAPACK's general eig (utils.py:1668) may order / sign its vectors differently for such inputs."""
    def fn(test_xs, X, Y):
        invK = onp.precomputeInvKs(np.asarray(l).reshape(1, -1), np.array([sigma]), np.array([sigma0]), X)
        _, cov = onp.compute_mean_var_f(np.asarray(test_xs), X, Y, l, sigma, sigma0, invK[0], fullcov=True)
        e, v = np.linalg.eigh(0.5 * (cov + cov.T))
        v = v * np.sign(v[np.argmax(np.abs(v), axis=0), np.arange(v.shape[1])])
        return v * np.sqrt(np.clip(e, 0.0, np.inf))
    return fn


@pytest.mark.parametrize("q", [1, 2])
def test_batch_bo_trajectory_matches_oracle_backend(golden, monkeypatch, q):
    """C1 (script_batch_branin.sh shapes, scaled: nmax 8, nfeature 64, ntrain 25): the device loop and the same loop
    on the numpy oracle make the same discrete decisions and follow the same trajectory (Adam paths agree to rounding:
    1e-5 on the query points)."""
    from tes_b200 import bo_batch
    import tes_b200.criteria.evaluate_batch_sample_mp as ebs
    f, xs, l, sigma, sigma0 = _branin(golden)
    args = dict(criterion="sftl", mode="sample", nquery=2, nrun=1, ninit=3, nmax=8, nfeature=64, nsto=2, ntrain=25,
                nysample=4, batchsize=q, ntestsample=300, seed=1,
                transform_fn=_stable_factor(l, sigma, sigma0))
    out = bo_batch.run(f, xs, 0.0, 1.0, l, sigma, sigma0, **args)
    ob = _OracleBackend()
    for name in ("precomputeInvKs", "compute_mean_f", "compute_mean_var_f", "get_testidxs_stats"):
        monkeypatch.setattr(bo_batch.utils, name, getattr(ob, name))
    monkeypatch.setattr(bo_batch.optfunc, "draw_random_init_weights_features_np", ob.draw_random_init_weights_features_np)
    monkeypatch.setattr(bo_batch.utils_for_continuous, "sample_xmaxs_fmaxs", ob.sample_xmaxs_fmaxs)
    monkeypatch.setattr(ebs, "mp", ob.batch_sample_mp)
    monkeypatch.setattr(bo_batch, "dev", lambda a: a)
    monkeypatch.setattr(bo_batch.evaluate_batch_mp_mod(), "get_pNK_test_obs",
                        lambda ls, sg, s0, nh, X, tx: onp.get_pNK_test_obs(ls, sg, s0, nh, X, tx))
    ref = bo_batch.run(f, xs, 0.0, 1.0, l, sigma, sigma0, **args)
    np.testing.assert_array_equal(out["ntest"], ref["ntest"])
    assert (ref["ntest"] >= 2).any() and np.isfinite(ref["crit_max"]).any()
    # Adam on finite-difference gradients amplifies the 1e-13 device/numpy differences of the criterion when the
    # gradient is small (m / sqrt(v)): the paths agree to ~1e-7, the discrete decisions exactly
    np.testing.assert_allclose(out["xx"], ref["xx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["guess_xx"], ref["guess_xx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["crit_max"], ref["crit_max"], rtol=1e-5, atol=1e-8)
    assert out["xx"].shape == (1, 3 + 2 * q, 2)


@pytest.mark.parametrize("criterion,mode,det", [("sftl", "sample", 0), ("ftl", "empirical", 0), ("ftl", "ep", 1)])
def test_sequential_bo_trajectory_matches_oracle_backend(golden, monkeypatch, criterion, mode, det):
    """bo.py loop (tes_b200.bo.run): de-duplicated maximisers, evaluate_sample_mp / evaluate_mp as the criterion, all
    maximisers as start points of its Adam refinement; device run vs the same loop on the numpy oracle."""
    from tes_b200 import bo, bo_batch
    import tes_b200.criteria.evaluate_mp as em
    import tes_b200.criteria.evaluate_sample_mp as es
    f, xs, l, sigma, sigma0 = _branin(golden)
    args = dict(criterion=criterion, mode=mode, deterministic=det, nquery=2, nrun=1, ninit=3, nmax=8, nfeature=64, nsto=2,
                ntrain=20, nysample=4, ntestsample=300, seed=2, transform_fn=_stable_factor(l, sigma, sigma0))
    out = bo.run(f, xs, 0.0, 1.0, l, sigma, sigma0, **args)
    ob = _OracleBackend()
    for name in ("precomputeInvKs", "compute_mean_f", "compute_mean_var_f", "get_testidxs_stats", "remove_duplicates_np"):
        monkeypatch.setattr(bo_batch.utils, name, getattr(ob, name))
    monkeypatch.setattr(bo_batch.optfunc, "draw_random_init_weights_features_np", ob.draw_random_init_weights_features_np)
    monkeypatch.setattr(bo_batch.utils_for_continuous, "sample_xmaxs_fmaxs", ob.sample_xmaxs_fmaxs)
    monkeypatch.setattr(es, "mp", ob.sample_mp)
    monkeypatch.setattr(em, "mp", ob.ep_mp)
    monkeypatch.setattr(bo_batch, "dev", lambda a: a)
    monkeypatch.setattr(bo_batch.evaluate_batch_mp_mod(), "get_pNK_test_obs",
                        lambda ls, sg, s0, nh, X, tx: onp.get_pNK_test_obs(ls, sg, s0, nh, X, tx))
    ref = bo.run(f, xs, 0.0, 1.0, l, sigma, sigma0, **args)
    np.testing.assert_array_equal(out["ntest"], ref["ntest"])
    assert (ref["ntest"] >= 2).any() and np.isfinite(ref["crit_max"]).any()
    np.testing.assert_allclose(out["xx"], ref["xx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["guess_xx"], ref["guess_xx"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["crit_max"], ref["crit_max"], rtol=1e-5, atol=1e-8)
