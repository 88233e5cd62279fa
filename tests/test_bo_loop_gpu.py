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
