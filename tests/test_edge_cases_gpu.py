"""Edge cases of the hot path through the reference-named API: candidate counts around the internal chunk size
(148 x 128 = 18 944 rows per pass), a single candidate, a lone trusted maximizer (the drivers then skip the criterion,
bo_discrete.py:557-559), S = 1, k = N, and empty candidate sets."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import tes_oracle_np as onp
from tests._problems import HART6, make_problem

pytestmark = pytest.mark.gpu
CHUNK = 148 * 128


@pytest.mark.parametrize("nx", [1, CHUNK - 1, CHUNK, CHUNK + 1, 2 * CHUNK + 7])
def test_tes_ep_chunk_boundaries(nx):
    """Deterministic TES_ep on nx candidates = the same values as scoring them in two arbitrary pieces, and the first /
    last rows agree with the oracle (the oracle on all 38k rows would take minutes)."""
    from tes_b200.criteria import evaluate_mp
    pr = make_problem(seed=5, d=6, n=10, T=7, nx=nx, hyp=HART6, mode="empirical", nsample=400)
    args = (pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 6)
    tail = (pr["n"], 1, 10, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"])
    acq = evaluate_mp.mp(pr["x"], *args, nx, *tail)
    assert acq.shape == (nx,) and np.all(np.isfinite(acq))
    cut = max(1, nx // 3)
    parts = [evaluate_mp.mp(pr["x"][:cut], *args, cut, *tail)]
    if nx > cut:
        parts.append(evaluate_mp.mp(pr["x"][cut:], *args, nx - cut, *tail))
    np.testing.assert_array_equal(np.concatenate(parts), acq)
    sel = np.unique(np.concatenate([np.arange(min(nx, 5)), np.arange(max(0, nx - 5), nx)]))
    ref = onp.evaluate_mp(pr["x"][sel], *args, len(sel), *tail)
    np.testing.assert_allclose(acq[sel], ref, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("nx", [1, CHUNK + 1])
def test_tes_sp_chunk_boundaries(nx):
    from tes_b200.criteria import evaluate_sample_mp
    pr = make_problem(seed=6, d=2, n=5, T=5, nx=nx, mode="sample", nsample=120)
    args = (pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2)
    tail = (pr["n"], 1, 3, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"])
    kw = dict(niteration=1, seed=12, n_global=nx + 10, x_offset=4)
    acq = evaluate_sample_mp.mp(pr["x"], *args, nx, *tail, **kw)
    assert acq.shape == (nx,) and np.all(np.isfinite(acq))
    # the draws are keyed on the GLOBAL candidate index: any split gives the same numbers
    cut = max(1, nx // 2)
    a = evaluate_sample_mp.mp(pr["x"][:cut], *args, cut, *tail, **kw)
    np.testing.assert_array_equal(a, acq[:cut])
    if nx > cut:
        kw2 = dict(kw, x_offset=4 + cut)
        b = evaluate_sample_mp.mp(pr["x"][cut:], *args, nx - cut, *tail, **kw2)
        np.testing.assert_array_equal(b, acq[cut:])


def test_lone_maximizer_is_queried_directly():
    """All function samples peak at the same candidate -> one unique maximizer -> the step returns it without scoring
    the criterion (bo_discrete.py:557-559), on both paths."""
    from oracle import pipeline_np
    from tes_b200 import acquisition
    rs = np.random.RandomState(0)
    d, S, F = 2, 4, 32
    cand = rs.rand(50, d)
    W = np.zeros((S, F, d))                      # constant function samples except for a bump at candidate 17
    b = np.zeros((S, F, 1))
    theta = np.zeros((S, F, 1))
    W[:, 0, :] = 1e-3
    theta[:, 0, 0] = 1.0
    b[:, 0, 0] = -1e-3 * cand[17].sum()          # cos(w.x + b) is maximal where w.x + b = 0
    X, Y = rs.rand(3, d), rs.randn(3, 1)
    l, sigma, sigma0 = np.array([5.0, 7.0]), 1.0, 1e-3
    ref = pipeline_np.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b)
    out = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b)
    assert ref["n_unique_maximizers"] == out["n_unique_maximizers"] == 1
    assert out["query_idx"] == ref["query_idx"] == 17
    np.testing.assert_array_equal(out["query_x"], cand[17])


def test_tiny_and_degenerate_top_k():
    from tes_b200 import ops
    rs = np.random.RandomState(1)
    d, F = 3, 40
    W, b, theta = rs.randn(1, F, d), rs.rand(1, F, 1), rs.randn(1, F, 1)
    # one candidate, one sample
    cand = rs.rand(1, d)
    idx, val = ops.rff_topk(cand, W, b, theta, 0.7, k=1)
    ridx, rval = co.rff_topk(cand, W, b, theta, 0.7, 1)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(val.cpu().numpy(), rval)
    # k = N: the whole candidate set, sorted by (value desc, index asc)
    cand = rs.rand(5, d)
    idx, val = ops.rff_topk(cand, W, b, theta, 0.7, k=5)
    ridx, rval = co.rff_topk(cand, W, b, theta, 0.7, 5)
    assert np.array_equal(idx.cpu().numpy(), ridx) and np.array_equal(val.cpu().numpy(), rval)
    # duplicated candidates: exact ties resolve to the lower index
    cand = np.repeat(rs.rand(1, d), 6, axis=0)
    idx, _ = ops.rff_topk(cand, W, b, theta, 0.7, k=3)
    assert idx.cpu().numpy().tolist() == [[0, 1, 2]]


def test_empty_candidate_sets():
    import torch
    from tes_b200 import ops
    from tes_b200.criteria import evaluate_mp
    pr = make_problem(seed=7, d=2, n=5, T=5, nx=3, mode="empirical", nsample=200)
    acq = evaluate_mp.mp(np.zeros((0, 2)), pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, 0, pr["n"], 1, 10,
                         pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"])
    assert acq.shape == (0,)
    k = ops.se_ard_cross(np.zeros((0, 2)), pr["X"], pr["l"], pr["sigma"])
    assert tuple(k.shape) == (0, 5)
    inv = ops.batched_chol2inv(torch.zeros((0, 4, 4), dtype=torch.float64, device="cuda"))
    assert tuple(inv.shape) == (0, 4, 4)


def test_pnk_inverse_falls_back_to_general_inverse_when_not_positive_definite():
    """pNK has no jitter on its test block; near-duplicate test points make the Cholesky route
    fail where the reference's tf.linalg.inv (LU) still returns an inverse."""
    from tes_b200 import ops
    from tes_b200.criteria import evaluate_mp
    rs = np.random.RandomState(3)
    d, n, T = 2, 6, 5
    X = rs.rand(n, d)
    test_xs = rs.rand(T, d)
    test_xs[3] = test_xs[1]                       # an exact duplicate: the test block is singular to rounding
    ls, sigmas, sigma0s = np.array([[25.5866, 6.3735]]), np.array([0.589]), np.array([1e-4])
    pNK, inv = evaluate_mp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)
    ref = onp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)[0]
    np.testing.assert_allclose(pNK, ref, rtol=1e-12, atol=1e-14)
    assert np.all(np.isfinite(inv) | np.isinf(inv))          # LU of a singular matrix: inf / huge, never silently NaN-free garbage
    # a merely ill-conditioned (not singular) case must agree with LAPACK's LU inverse through the residual
    test_xs[3] = test_xs[1] + 3e-6
    pNK, inv = evaluate_mp.get_pNK_test_obs(ls, sigmas, sigma0s, 1, X, test_xs)
    gj = ops.gj_inverse(pNK[0]).cpu().numpy()
    lu = np.linalg.inv(pNK[0])
    scale = np.abs(lu).max()
    np.testing.assert_allclose(gj, lu, rtol=0, atol=1e-4 * scale)
    assert np.abs(pNK[0] @ gj - np.eye(T + n)).max() < 1e-4


def test_checked_cholesky_raises_like_tf_cholesky():
    from tes_b200 import utils
    from tes_b200._lib import TesError
    A = np.array([[1.0, 2.0], [2.0, 1.0]])        # indefinite
    with pytest.raises(TesError):
        utils.chol2inv(A)
    with pytest.raises(TesError):
        utils.multichol2inv(np.stack([np.eye(2), A]))


def test_screen_bound_with_wide_dynamic_range_and_indefinite_covariance():
    """the screen's error bound carries the fp16-subnormal absolute term (entries of Sigma_j / M far below the
    column / row maxima) and is +inf for a covariance that violates |S_ab| <= sqrt(S_aa S_bb)."""
    from tes_b200 import ops
    rs = np.random.RandomState(11)
    d, n, T, N, Sp = 3, 8, 12, 40000, 6
    X, Y = rs.rand(n, d), rs.randn(n, 1) * 0.3
    test_xs = rs.rand(T, d)
    x = rs.rand(N, d)
    l, sigma, sigma0 = np.array([9.0, 4.0, 2.0]), 1.1, 1e-3
    ls, sg, s0 = l.reshape(1, d), np.array([sigma]), np.array([sigma0])
    invK = onp.precomputeInvKs(ls, sg, s0, X)
    _, invpNK = onp.get_pNK_test_obs(ls, sg, s0, 1, X, test_xs)
    p = np.full(Sp, 1.0 / Sp)
    # covariances whose entries span > 1e7 in magnitude: D A D with D = diag(10^-4 ... 10^0)
    covs = []
    for j in range(Sp):
        A = rs.randn(T, T)
        A = A @ A.T / T + 0.1 * np.eye(T)
        D = np.diag(10.0 ** rs.uniform(-4, 0, size=T))
        covs.append(D @ A @ D)
    covs = np.stack(covs)
    acq_t, eps = ops.ep_acq_screen(x, test_xs, X, Y, l, sigma, sigma0, invpNK[0], invK[0], p, covs)
    ref = ops.ep_acq(ops.EP_DETERMINISTIC, x, test_xs, X, Y, l, sigma, sigma0, invpNK[0], invK[0], p,
                     np.zeros((Sp, T)), covs)
    err = (acq_t - ref).abs()
    assert bool(torch_all(err <= eps)), float((err - eps).max())
    assert float(eps.max()) < 1e-2                       # ... and the bound is still useful
    # an indefinite "covariance" (EP mode can produce one): the bound must give up, not lie
    covs_bad = covs.copy()
    covs_bad[2, 0, 1] = covs_bad[2, 1, 0] = 10.0 * np.sqrt(covs_bad[2, 0, 0] * covs_bad[2, 1, 1])
    _, eps_bad = ops.ep_acq_screen(x, test_xs, X, Y, l, sigma, sigma0, invpNK[0], invK[0], p, covs_bad)
    assert bool(torch_all(eps_bad == float("inf")))


def torch_all(t):
    import torch
    return torch.all(t)
