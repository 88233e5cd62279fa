"""End-to-end parity of one TES acquisition step (stages A-D) on the reference's own configurations:
  C1  Branin 2-D, 20x20 grid, S=30, F=300 (script_batch_branin.sh), TES_sp (sftl / sample) and TES_ep
  C2  Brooms-Barn pH 2-D, 20x20 grid, shipped hyper-parameters, TES_ep with mode ep and empirical
Selections (trusted maximizers, query index) must be IDENTICAL to the oracle's; acquisition values
within 1e-6 relative (+1e-9 absolute)."""
import numpy as np
import pytest

from oracle import pipeline_np
from oracle import tes_oracle_np as onp

pytestmark = pytest.mark.gpu


def _draw(rs, S, F, X, Y, l, sigma, sigma0):
    d = X.shape[1]
    return onp.draw_random_init_weights_features(d, S, F, X, Y, l, sigma, sigma0, rs.randn(S, F, d),
                                                 rs.rand(S, F, 1), rs.randn(S, F, 1))


def _transform(cand, X, Y, l, sigma, sigma0, theta, W, b, t_max=None):
    """The colouring factor A of utils.py:1668-1670, computed once (oracle side) and handed to BOTH
    paths: np.linalg.eig's column order/signs are not stable under the 1e-13 differences between the CPU
    and the CUDA posterior covariance, and A z is part of the host-supplied draw."""
    probe = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, criterion="ftl", mode="empirical",
                                 ntestsample=10, t_max=t_max)
    return onp.mvn_transform_matrix(probe["test_cov"])


def _compare(out, ref, acq_tol=(1e-6, 1e-9)):
    assert np.array_equal(out["argmax_idx"].cpu().numpy(), ref["argmax_idx"])
    assert out["n_unique_maximizers"] == ref["n_unique_maximizers"]
    assert (out["T"], out["Sp"]) == (ref["T"], ref["Sp"])
    np.testing.assert_allclose(out["acq"].cpu().numpy(), ref["acq"], rtol=acq_tol[0], atol=acq_tol[1])
    assert out["query_idx"] == ref["query_idx"]
    np.testing.assert_array_equal(out["query_x"], ref["query_x"])
    assert abs(out["query_val"] - ref["query_val"]) <= acq_tol[0] * abs(ref["query_val"]) + acq_tol[1]


@pytest.mark.parametrize("criterion,mode,det", [("ftl", "empirical", True), ("ftl", "empirical", False),
                                                ("ftl", "ep", True)])
def test_c1_branin_step(golden, criterion, mode, det):
    from tes_b200 import acquisition
    g = golden("kat_objectives")
    cand, l = g["branin_xs"], g["branin_l"]
    sigma, sigma0 = float(g["branin_sigma"]), float(g["branin_sigma0"])
    rs = np.random.RandomState(1)
    X = rs.rand(4, 2)
    idx_obs = [np.argmin(((cand - x) ** 2).sum(1)) for x in X]
    Y = g["branin_f_on_xs"][idx_obs].reshape(-1, 1) + rs.randn(4, 1) * np.sqrt(sigma0)
    theta, W, b = _draw(rs, 30, 300, X, Y, l, sigma, sigma0)
    kw = dict(criterion=criterion, mode=mode, deterministic=det, ntestsample=1000, nysample=6, niteration=2,
              seed=3, return_acq=True, transform=_transform(cand, X, Y, l, sigma, sigma0, theta, W, b))
    z = None
    if not det:
        probe = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, criterion="ftl", mode=mode,
                                     deterministic=True, ntestsample=1000, seed=3, transform=kw["transform"])
        z = np.random.RandomState(9).normal(size=(1, 2, 6, probe["Sp"], 400))
    ref = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, z=z, **kw)
    out = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, z=z, **kw)
    _compare(out, ref)


def test_c1_branin_tes_sp_subset(golden):
    """TES_sp (criterion sftl, mode sample) on the Branin problem.  Stage A runs on the full 20x20
    grid; the criterion is then scored on a 16-candidate subset with host-supplied draws (the O(S'^2 K)
    oracle tensor for all 400 candidates does not fit a CPU test)."""
    from tes_b200 import acquisition, ops
    from tes_b200.criteria import evaluate_sample_mp
    g = golden("kat_objectives")
    cand, l = g["branin_xs"], g["branin_l"]
    sigma, sigma0 = float(g["branin_sigma"]), float(g["branin_sigma0"])
    rs = np.random.RandomState(2)
    X = rs.rand(5, 2)
    Y = rs.randn(5, 1) * 0.3
    theta, W, b = _draw(rs, 30, 300, X, Y, l, sigma, sigma0)
    # stages A-C through both paths
    idx, val, pts = acquisition.trusted_maximizers(cand, W, b, theta, sigma)
    from oracle import c_oracle as co
    ridx, rval = co.rff_topk(cand, W, b, theta, sigma, 1)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    test_xs = acquisition.select_test_points(pts[:, 0], 1e-3).cpu().numpy()
    np.testing.assert_array_equal(test_xs, pipeline_np.select_test_points(cand[ridx[:, 0]], 1e-3))
    ls, sg, s0 = l.reshape(1, 2), np.array([sigma]), np.array([sigma0])
    invK = onp.precomputeInvKs(ls, sg, s0, X)
    mean_t, cov_t = onp.compute_mean_var_f(test_xs, X, Y, l, sigma, sigma0, invK[0], fullcov=True)
    T0 = test_xs.shape[0]
    zj = rs.normal(size=(1, 1000, T0, 1))
    tk, mp_, mk, ck, v0, v1 = onp.get_testidxs_stats(1, test_xs, mean_t.reshape(1, -1), cov_t[None], "sample",
                                                     1000, 2, z=zj)
    _, invpNK = onp.get_pNK_test_obs(ls, sg, s0, 1, X, tk)
    sub = cand[::25]
    nx, Sp, K = sub.shape[0], int((mp_[0] > 0).sum()), v0.shape[-1]
    z = rs.normal(size=(1, 2, 4, Sp, nx, K))
    ref = onp.evaluate_sample_mp(sub, ls, sg, s0, X, Y, 2, nx, 5, 1, 4, tk, mp_, v0, v1, invpNK, niteration=2, z=z)
    got = evaluate_sample_mp.mp(sub, ls, sg, s0, X, Y, 2, nx, 5, 1, 4, tk, mp_, v0, v1, invpNK, niteration=2, z=z)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-9)
    assert int(np.argmax(got)) == int(np.argmax(ref))


@pytest.mark.parametrize("mode,S", [("ep", 5), ("empirical", 5), ("ep", 32), ("empirical", 32)])
def test_c2_ph_discrete_step(golden, mode, S):
    """bo_discrete.py defaults (S=5) and the widened S=32, pH objective, shipped hyper-parameters."""
    from tes_b200 import acquisition
    g = golden("kat_objectives")
    hyp = g["pH_hyp"]
    sigma, l, sigma0 = float(hyp[0]), hyp[1:3], float(hyp[3])
    cand, f = g["pH_xs"], g["pH_f_on_xs"]
    rs = np.random.RandomState(S)
    obs = rs.choice(400, 12, replace=False)       # observations sampled from the grid (bo_discrete.py:938-941)
    X = cand[obs]
    Y = (f[obs] + rs.randn(12) * np.sqrt(sigma0)).reshape(-1, 1)
    theta, W, b = _draw(rs, S, 300, X, Y, l, sigma, sigma0)
    kw = dict(criterion="ftl", mode=mode, deterministic=True, ntestsample=1000, seed=5, return_acq=True,
              transform=_transform(cand, X, Y, l, sigma, sigma0, theta, W, b))
    ref = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, **kw)
    out = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, **kw)
    _compare(out, ref)
    # stochastic estimator (the reference's default, --deterministic 0) with host draws, I=50, Y=10 scaled down
    z = rs.normal(size=(1, 5, 10, ref["Sp"], 400))
    kw.update(deterministic=False, niteration=5, nysample=10)
    ref = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, z=z, **kw)
    out = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, z=z, **kw)
    _compare(out, ref)


@pytest.mark.parametrize("criterion,det,q", [("ftl", False, 1), ("sftl", True, 1), ("sftl", True, 3), ("sftl", True, 8)])
def test_step_with_counter_based_draws(criterion, det, q):
    """Monte-Carlo criteria with the counter-based draws of include/tes_spec.h (no host z): the oracle pipeline
    regenerates the same normals from (seed, iteration, stream, element), so selections are identical and values
    agree to 1e-6.  q > 1: batch TES_sp (evaluate_batch_sample_mp.py) on every q consecutive candidates (C4 shape)."""
    from tes_b200 import acquisition
    d, n, S, F, N = 3, 6, 12, 64, 48
    rs = np.random.RandomState(40 + q)
    l, sigma, sigma0 = np.array([9.0, 5.0, 7.0]), 1.3, 1e-3
    cand, X = rs.rand(N, d), rs.rand(n, d)
    Y = rs.randn(n, 1) * 0.4
    theta, W, b = _draw(rs, S, F, X, Y, l, sigma, sigma0)
    kw = dict(criterion=criterion, mode="empirical", deterministic=det, ntestsample=300, nysample=3, niteration=2,
              seed=77, return_acq=True, q=q, t_max=4, transform=_transform(cand, X, Y, l, sigma, sigma0, theta, W, b, 4))
    ref = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, **kw)
    out = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, **kw)
    assert ref["T"] >= 2 and out["acq"].shape[0] == N // q
    _compare(out, ref)


def test_step_from_a_pinned_host_candidate_array_overlaps_the_upload_and_gives_the_same_step():
    """tes_step(draws=...) uploads a pinned host candidate tensor on a side stream under the weight draw (stage A0 does not
    read the candidates): same outputs as with the candidates already on the device, repeatedly (allocator reuse)."""
    import torch
    from tes_b200 import acquisition
    rs = np.random.RandomState(4)
    d, n, S, F, N = 3, 9, 24, 96, 40_000
    l, sigma, sigma0 = np.array([5.0, 2.0, 8.0]), 1.2, 1e-3
    cand, X, Y = rs.rand(N, d), rs.rand(n, d), rs.randn(n, 1) * 0.4
    draws = tuple(torch.from_numpy(a).cuda() for a in (rs.randn(S, F, d), rs.rand(S, F, 1), rs.randn(S, F, 1)))
    kw = dict(criterion="ftl", mode="empirical", deterministic=True, ntestsample=2000, seed=3, t_max=32, draws=draws,
              return_acq=True)
    ref = acquisition.tes_step(torch.from_numpy(cand).cuda(), X, Y, l, sigma, sigma0, None, None, None, **kw)
    host = torch.from_numpy(cand).pin_memory()
    for _ in range(3):
        out = acquisition.tes_step(host, X, Y, l, sigma, sigma0, None, None, None, **kw)
        assert out["query_idx"] == ref["query_idx"] and out["query_val"] == ref["query_val"]
        assert torch.equal(out["argmax_idx"], ref["argmax_idx"]) and torch.equal(out["acq"], ref["acq"])
