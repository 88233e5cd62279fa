"""Stage B parity (SE-ARD kernels, pNK, batched chol2inv, posterior statistics) vs the numpy oracle.
Tolerance: 1e-6 relative is the north-star bar for acquisition values; the building blocks are held
to much tighter bounds, stated per test."""
import numpy as np
import pytest

from oracle import tes_oracle_np as onp
from tests._problems import BRANIN, HART6, PH, make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_se_ard_vs_reference_golden(golden, tag):
    from tes_b200 import ops
    g = golden("se_ard")
    K = ops.se_ard_cross(g[tag + "_A"], g[tag + "_B"], g[tag + "_l"], float(g[tag + "_sigma"])).cpu().numpy()
    np.testing.assert_allclose(K, g[tag + "_Knm"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("N,m,d", [(1, 1, 1), (33, 31, 2), (1000, 189, 6), (257, 65, 10), (70, 40, 16), (50, 20, 17)])
def test_se_ard_tiled_kernel_is_bit_identical_to_the_per_element_kernel(N, m, d, monkeypatch):
    """se_ard_tile_kernel (scaled coordinates and norms staged once per 32 x 32 tile) performs the per-element kernel's
    operations in the same order: identical bits, and both within 1e-13 of the numpy oracle (utils.py:836-859);
    d = 17 > SE_MAXD takes the per-element kernel either way."""
    from tes_b200 import ops
    rs = np.random.RandomState(N + m + d)
    A, B = rs.rand(N, d) * 3 - 1, rs.rand(m, d) * 3 - 1
    B[: min(m, N) // 2] = A[: min(m, N) // 2]              # exact duplicates: the cancellation case of the expansion
    l = rs.rand(d) * 4 + 0.1
    monkeypatch.setenv("TES_SE_ARD_TILED", "1")
    K1 = ops.se_ard_cross(A, B, l, 1.37).cpu().numpy()
    monkeypatch.setenv("TES_SE_ARD_TILED", "0")
    K0 = ops.se_ard_cross(A, B, l, 1.37).cpu().numpy()
    np.testing.assert_array_equal(K1, K0)
    np.testing.assert_allclose(K1, onp.computeKnm(A, B, l, 1.37), rtol=1e-13, atol=1e-15)


def test_brooms_barn_known_answer_through_cuda(golden):
    """The reference's only KAT (functions.py:397-398): pH objective maximum 0.73425606, evaluated
    with the CUDA SE-ARD kernel + CUDA Cholesky inverse + CUDA posterior mean."""
    import torch
    from tes_b200 import ops
    g = golden("kat_objectives")
    X, Y, hyp = g["pH_X"], g["pH_Y"], g["pH_hyp"]
    sigma, l, sigma0 = hyp[0], hyp[1:3], hyp[3]
    NK = ops.se_ard_cross(X, X, l, sigma) + torch.eye(X.shape[0], dtype=torch.float64, device="cuda") * sigma0
    NKinv = ops.batched_chol2inv(NK)
    mean, var = ops.gp_mean_var(g["pH_maximizer"].reshape(1, 2), X, Y, l, sigma, NKinv)
    assert abs(float(mean[0]) - 0.73425606) < 5e-9
    mean_xs, _ = ops.gp_mean_var(g["pH_xs"], X, Y, l, sigma, NKinv)
    np.testing.assert_allclose(mean_xs.cpu().numpy(), g["pH_f_on_xs"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("batch,m", [(1, 1), (3, 5), (1, 64), (4, 65), (2, 130), (1, 200), (2, 257)])
def test_batched_chol2inv(batch, m):
    from tes_b200 import ops
    rs = np.random.RandomState(m)
    d = 3
    mats = []
    for _ in range(batch):
        Xp = rs.rand(m, d)
        mats.append(onp.computeKmm(Xp, np.array([3.0, 5.0, 2.0]), 1.3) + np.eye(m) * 1e-2)
    A = np.stack(mats)
    inv, info = ops.batched_chol2inv(A, return_info=True)
    assert int(info.abs().sum()) == 0
    ref = onp.multichol2inv(A)
    inv = inv.cpu().numpy()
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
    # cond ~ 1e4..1e5: agreement to ~1e-10 of the largest entry
    assert np.max(np.abs(inv - ref) / scale) < 1e-9
    L = ops.batched_potrf(A).cpu().numpy()
    np.testing.assert_allclose(L, np.linalg.cholesky(A), rtol=1e-10, atol=1e-12)
    # identity check
    eye = np.eye(m)[None]
    assert np.max(np.abs(A @ inv - eye)) < 1e-8


def test_chol2inv_flags_non_spd():
    from tes_b200 import ops
    A = np.eye(70)
    A[40, 40] = -1.0
    inv, info = ops.batched_chol2inv(A[None], return_info=True)
    assert int(info[0]) == 41
    assert not np.isfinite(inv.cpu().numpy()).all()


@pytest.mark.parametrize("hyp,d,n,T", [(BRANIN, 2, 5, 6), (PH, 2, 30, 5), (HART6, 6, 40, 24)])
def test_pnk_and_posterior_stats(hyp, d, n, T):
    from tes_b200 import ops
    pr = make_problem(seed=n, d=d, n=n, T=T, nx=300, hyp=hyp)
    pNK_ref, inv_ref = onp.get_pNK_test_obs(pr["ls"], pr["sigmas"], pr["sigma0s"], 1, pr["X"], pr["test_xs"])
    pNK = ops.pnk(pr["test_xs"], pr["X"], pr["l"], pr["sigma"], pr["sigma0"])
    np.testing.assert_allclose(pNK.cpu().numpy(), pNK_ref[0], rtol=1e-13, atol=1e-15)
    inv = ops.batched_chol2inv(pNK).cpu().numpy()
    assert np.max(np.abs(inv - inv_ref[0])) / np.abs(inv_ref[0]).max() < 1e-7  # LU vs Cholesky, cond-limited
    # posterior stats on the oracle's own inverse (identical input)
    M, b, Kq = ops.posterior_stats(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["invpNK"][0])
    Tk = pr["T"]
    k = onp.computeKnm(pr["x"], np.concatenate([pr["test_xs"], pr["X"]]), pr["l"], pr["sigma"])
    Mr = k @ pr["invpNK"][0][:, :Tk]
    br = (k @ pr["invpNK"][0][:, Tk:] @ pr["Y"]).ravel()
    Kqr = pr["sigma"] - np.sum((k @ pr["invpNK"][0]) * k, axis=1)
    scale = np.abs(pr["invpNK"][0]).max() * pr["sigma"]
    assert np.max(np.abs(M.cpu().numpy() - Mr)) < 1e-12 * scale
    assert np.max(np.abs(b.cpu().numpy() - br)) < 1e-12 * scale
    assert np.max(np.abs(Kq.cpu().numpy() - Kqr)) < 1e-12 * scale * pr["sigma"] * (Tk + n)
    mean, var = ops.gp_mean_var(pr["x"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["invK"][0])
    mr, vr = onp.compute_mean_var_f(pr["x"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invK"][0])
    s2 = np.abs(pr["invK"][0]).max() * pr["sigma"] ** 2 * n
    assert np.max(np.abs(mean.cpu().numpy() - mr)) < 1e-12 * s2
    assert np.max(np.abs(var.cpu().numpy() - vr)) < 1e-12 * s2


@pytest.mark.parametrize("M,N,K,transb", [(1, 1, 1, False), (130, 70, 33, False), (257, 64, 16, True),
                                          (1000, 193, 192, False), (64, 300, 500, True)])
def test_gemm(M, N, K, transb):
    from tes_b200 import ops
    rs = np.random.RandomState(M + N + K)
    A = rs.randn(M, K)
    B = rs.randn(N, K) if transb else rs.randn(K, N)
    C = ops.gemm(A, B, transb=transb).cpu().numpy()
    ref = A @ (B.T if transb else B)
    np.testing.assert_allclose(C, ref, rtol=0, atol=1e-13 * K)


def test_argmax_first_maximum_and_nan():
    from tes_b200 import ops
    rs = np.random.RandomState(0)
    v = rs.randn(100_003)
    v[[777, 50_000, 99_999]] = 9.0
    i, val = ops.argmax(v)
    assert int(i) == 777 and float(val) == 9.0 == v[int(np.argmax(v))]
    v[60_000] = np.nan
    v[70_000] = np.nan
    i, val = ops.argmax(v)
    assert int(i) == int(np.argmax(v)) == 60_000 and np.isnan(float(val))
    v[100] = np.inf                      # np.argmax ranks NaN above +inf: still the first NaN
    i, val = ops.argmax(v)
    assert int(i) == int(np.argmax(v)) == 60_000 and np.isnan(float(val))
    v[[60_000, 70_000]] = 0.0
    i, val = ops.argmax(v)
    assert int(i) == 100 and float(val) == np.inf
    i, val = ops.argmax(np.array([3.0]))
    assert int(i) == 0


def test_discrete_maximizer_sampler(golden):
    """bo_discrete.py:359-400 on the pH grid: identical arg-max indices for identical z."""
    from tes_b200 import bo_discrete
    g = golden("kat_objectives")
    hyp = g["pH_hyp"]
    sigma, l, sigma0 = float(hyp[0]), hyp[1:3], float(hyp[3])
    xs, f = g["pH_xs"], g["pH_f_on_xs"]
    rs = np.random.RandomState(3)
    obs = rs.choice(400, 10, replace=False)
    X, Y = xs[obs], (f[obs] + rs.randn(10) * np.sqrt(sigma0)).reshape(-1, 1)
    z = rs.randn(7, 400)
    mx, fs, idx, mean, cov = bo_discrete.sample_maximizers(xs, X, Y, l, sigma, sigma0, z)
    ridx, rfs = onp.discrete_maximizer_samples(xs, X, Y, l, sigma, sigma0, z)
    np.testing.assert_array_equal(idx, ridx)
    # sqrtm's null-space part (|e| ~ 1e-17 -> sqrt ~ 3e-9, LAPACK-dependent singular vectors) bounds the agreement
    np.testing.assert_allclose(fs, rfs, rtol=0, atol=1e-7)
    np.testing.assert_array_equal(mx, xs[ridx])


@pytest.mark.parametrize("batch,m", [(3, 512), (2, 1024), (2, 449), (1, 2048), (2, 1150)])
def test_batched_chol2inv_large_property(batch, m):
    """C5 sizes: A A^-1 = I and symmetry (the oracle is too slow to be the checker at the full sweep)."""
    import torch
    from tes_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(m)
    Xp = torch.rand(batch, m, 6, dtype=torch.float64, device="cuda", generator=g)
    D = ((Xp[:, :, None, :] - Xp[:, None, :, :]) ** 2).sum(-1)
    A = 1.4 * torch.exp(-0.5 * 3.0 * D) + 1e-2 * torch.eye(m, dtype=torch.float64, device="cuda")
    inv, info = ops.batched_chol2inv(A, return_info=True)
    assert int(info.abs().sum()) == 0
    eye = torch.eye(m, dtype=torch.float64, device="cuda")
    assert float((A @ inv - eye).abs().max()) < 1e-7
    assert float((inv - inv.transpose(1, 2)).abs().max()) / float(inv.abs().max()) < 1e-12
    ref = np.linalg.inv(A[0].cpu().numpy())
    assert np.abs(inv[0].cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-8
