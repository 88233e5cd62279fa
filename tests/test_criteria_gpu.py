"""Stage D parity: the TES criteria (CUDA, through the C ABI) vs the numpy oracle on identical
host-supplied draws.  Bar (north_star): acquisition values within 1e-6 relative in fp64; we add an
absolute floor of 1e-9 for values near zero and, for the batch TES_ep criterion (which is ~0 by
construction), an absolute bound of 1e-12.  Selections (arg-max) must be identical."""
import numpy as np
import pytest

from oracle import tes_oracle_np as onp
from tests._problems import BRANIN, HART6, PH, make_problem

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-6, 1e-9


def _close(a, b):
    a = np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL)
    assert int(np.argmax(a)) == int(np.argmax(b))


def _ep_inputs(pr):
    p = pr["max_probs"][0]
    keep = p > 0
    return p[keep], pr["v0"][0][keep], pr["v1"][0][keep]


@pytest.mark.parametrize("hyp,d,n,T,nx", [(BRANIN, 2, 5, 6, 50), (PH, 2, 20, 5, 400), (HART6, 6, 30, 20, 200),
                                          (BRANIN, 2, 3, 40, 130)])
def test_tes_ep_deterministic_and_lite_and_moments(hyp, d, n, T, nx):
    from tes_b200 import ops
    pr = make_problem(seed=T, d=d, n=n, T=T, nx=nx, hyp=hyp, mode="empirical", nsample=600)
    p, pm, pc = _ep_inputs(pr)
    args = (pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
            pr["invK"][0], p, pm, pc)
    acq, mean, var = ops.ep_acq(ops.EP_DETERMINISTIC, *args, want_moments=True)
    qm, qv = onp.get_queried_f_stat_given_data_maxidx(pr["x"], pr["l"], pr["sigma"], pr["X"], pr["Y"],
                                                      pr["test_xs"], pr["invpNK"][0], pm, pc)
    np.testing.assert_allclose(mean.cpu().numpy().T, qm, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var.cpu().numpy().T, qv, rtol=1e-7, atol=1e-10)
    ref = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, nx, n, 1, 10,
                          pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"],
                          stochastic=False)
    _close(acq.cpu().numpy(), ref)
    lite = ops.ep_acq(ops.EP_LITE, *args)
    ref_l = onp.evaluate_mp_lite(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, nx, n, 1,
                                 pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"])
    _close(lite.cpu().numpy(), ref_l)


@pytest.mark.parametrize("hyp,d,n,T,nx,I,Yn", [(BRANIN, 2, 5, 6, 40, 3, 7), (PH, 2, 20, 5, 100, 5, 10),
                                               (HART6, 6, 30, 12, 60, 2, 4)])
def test_tes_ep_stochastic_host_draws(hyp, d, n, T, nx, I, Yn):
    from tes_b200 import ops
    pr = make_problem(seed=T + 1, d=d, n=n, T=T, nx=nx, hyp=hyp, mode="empirical", nsample=500)
    p, pm, pc = _ep_inputs(pr)
    Sp = p.shape[0]
    z = pr["rs"].normal(size=(1, I, Yn, Sp, nx))
    acq = ops.ep_acq(ops.EP_STOCHASTIC, pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"],
                     pr["sigma0"], pr["invpNK"][0], pr["invK"][0], p, pm, pc, niter=I, nysample=Yn, z=z[0])
    ref = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, nx, n, 1, Yn,
                          pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"],
                          stochastic=True, niteration=I, z=z)
    _close(acq.cpu().numpy(), ref)


def test_tes_ep_stochastic_philox_matches_oracle_generator():
    """Counter-based draws: the kernel's in-register normals equal the oracle generator's (to libm
    rounding), so feeding the oracle-generated tensor as host z reproduces the Philox run."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    from tes_b200.spec import STREAM_EP_Y
    pr = make_problem(seed=3, d=2, n=5, T=6, nx=33, hyp=BRANIN, mode="empirical", nsample=500)
    p, pm, pc = _ep_inputs(pr)
    Sp, nx, I, Yn, seed = p.shape[0], 33, 3, 5, 1234567
    args = (pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
            pr["invK"][0], p, pm, pc)
    a = ops.ep_acq(ops.EP_STOCHASTIC, *args, niter=I, nysample=Yn, seed=seed).cpu().numpy()
    z = np.stack([co.normals(seed, it, STREAM_EP_Y, 0, Yn * Sp * nx).reshape(Yn, Sp, nx) for it in range(I)])
    b = ops.ep_acq(ops.EP_STOCHASTIC, *args, niter=I, nysample=Yn, z=z).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    # sharding invariance: two shards with x_offset / n_global give the same numbers
    a1 = ops.ep_acq(ops.EP_STOCHASTIC, pr["x"][:10], *args[1:], niter=I, nysample=Yn, seed=seed, n_global=nx,
                    x_offset=0).cpu().numpy()
    a2 = ops.ep_acq(ops.EP_STOCHASTIC, pr["x"][10:], *args[1:], niter=I, nysample=Yn, seed=seed, n_global=nx,
                    x_offset=10).cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([a1, a2]), a)


@pytest.mark.parametrize("hyp,d,n,T,nx,I,Yn,ns", [(BRANIN, 2, 5, 6, 9, 2, 5, 200), (PH, 2, 12, 5, 6, 3, 4, 120),
                                                  (HART6, 6, 20, 10, 5, 1, 3, 300)])
def test_tes_sp_host_draws(hyp, d, n, T, nx, I, Yn, ns):
    from tes_b200 import ops
    pr = make_problem(seed=T + 2, d=d, n=n, T=T, nx=nx, hyp=hyp, mode="sample", nsample=ns)
    p = pr["max_probs"][0]
    keep = p > 0
    P, msk = pr["v0"][0][keep], pr["v1"][0][keep]
    Sp, K = P.shape[0], P.shape[2]
    z = pr["rs"].normal(size=(1, I, Yn, Sp, nx, K))
    acq = ops.sp_acq(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"],
                     pr["invpNK"][0], p[keep], P, msk, niter=I, nysample=Yn, z=z[0])
    ref = onp.evaluate_sample_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, nx, n, 1, Yn,
                                 pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"], niteration=I,
                                 z=z)
    _close(acq.cpu().numpy(), ref)


@pytest.mark.parametrize("q", [1, 2, 3, 8])
def test_tes_sp_batch_host_draws(q):
    from tes_b200 import ops
    d, n, T, nx, I, Yn = 2, 5, 6, 5, 2, 3
    pr = make_problem(seed=40 + q, d=d, n=n, T=T, nx=nx, hyp=BRANIN, mode="sample", nsample=150)
    p = pr["max_probs"][0]
    keep = p > 0
    P, msk = pr["v0"][0][keep], pr["v1"][0][keep]
    Sp, K = P.shape[0], P.shape[2]
    xb = pr["rs"].rand(nx, q, d)
    z = pr["rs"].normal(size=(1, I, Yn, Sp, nx, K, q))
    acq = ops.sp_acq(xb, pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                     p[keep], P, msk, niter=I, nysample=Yn, z=z[0])
    ref = onp.evaluate_batch_sample_mp(xb, pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, n, 1, Yn,
                                       pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"],
                                       niteration=I, z=z)
    _close(acq.cpu().numpy(), ref)


def test_batch_tes_ep_is_zero_like_the_reference():
    from tes_b200 import ops
    d, n, T, nx, q, I, Yn = 2, 5, 6, 4, 3, 2, 3
    pr = make_problem(seed=77, d=d, n=n, T=T, nx=nx, hyp=BRANIN, mode="empirical", nsample=300)
    p, pm, pc = _ep_inputs(pr)
    xb = pr["rs"].rand(nx, q, d)
    z = pr["rs"].normal(size=(1, I, Yn, p.shape[0], nx, q))
    ref = onp.evaluate_batch_mp(xb, pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, n, 1, Yn,
                                pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"],
                                niteration=I, z=z)
    acq = ops.batch_ep_acq(p, xb).cpu().numpy()
    assert np.max(np.abs(ref)) < 1e-12 and np.max(np.abs(acq - ref)) < 1e-12  # absolute bound


def test_negative_variance_propagates_nan():
    """no clipping before sqrt(var + sigma0) -- NaN reaches the output (driver retries)."""
    from tes_b200 import ops
    pr = make_problem(seed=5, d=2, n=5, T=6, nx=8, hyp=BRANIN, mode="empirical", nsample=300)
    p, pm, pc = _ep_inputs(pr)
    pc = pc.copy()
    pc[0] = -np.eye(pc.shape[1]) * 1e3  # forces var_0 + sigma0 < 0 somewhere
    acq = ops.ep_acq(ops.EP_DETERMINISTIC, pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"],
                     pr["sigma0"], pr["invpNK"][0], pr["invK"][0], p, pm, pc).cpu().numpy()
    pv1 = pr["v1"].copy()
    pv1[0][np.where(pr["max_probs"][0] > 0)[0][0]] = pc[0]
    ref = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, 8, 5, 1, 10,
                          pr["test_xs"], pr["max_probs"], pr["v0"], pv1, pr["invK"], pr["invpNK"])
    assert np.array_equal(np.isnan(acq), np.isnan(ref)) and np.isnan(ref).any()


def test_c4_shapes_batch_tes_sp_philox():
    """C4 shapes scaled down: d = 10, batch q = 8, T forced small, Philox draws.  The oracle generator
    (oracle/tes_oracle.c) reproduces the in-kernel normals, so feeding them as host z gives the same numbers;
    the numpy oracle then checks the value."""
    from oracle import c_oracle as co
    from tes_b200 import ops
    from tes_b200.spec import STREAM_BSP_Y
    d, n, T, nx, q, I, Yn, seed = 10, 16, 8, 6, 8, 1, 3, 424242
    hyp = dict(l=np.full(10, 1.0), sigma=2.0, sigma0=1e-3)       # functions.func_gp_prior(10, l=1.0, sigma=2.0)
    pr = make_problem(seed=4, d=d, n=n, T=T, nx=nx, hyp=hyp, mode="sample", nsample=200)
    p = pr["max_probs"][0]
    keep = p > 0
    P, msk = pr["v0"][0][keep], pr["v1"][0][keep]
    Sp, K = P.shape[0], P.shape[2]
    xb = pr["rs"].rand(nx, q, d)
    a = ops.sp_acq(xb, pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                   p[keep], P, msk, niter=I, nysample=Yn, seed=seed).cpu().numpy()
    z = np.stack([co.normals(seed, it, STREAM_BSP_Y, 0, Yn * Sp * nx * K * q).reshape(Yn, Sp, nx, K, q)
                  for it in range(I)])
    b = ops.sp_acq(xb, pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                   p[keep], P, msk, niter=I, nysample=Yn, z=z).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    ref = onp.evaluate_batch_sample_mp(xb, pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], d, n, 1, Yn,
                                       pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"],
                                       niteration=I, z=z[None])
    _close(b, ref)
    # sharding invariance of the Philox draws (global query index keys the stream)
    a1 = ops.sp_acq(xb[:2], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                    p[keep], P, msk, niter=I, nysample=Yn, seed=seed, n_global=nx, x_offset=0).cpu().numpy()
    a2 = ops.sp_acq(xb[2:], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                    p[keep], P, msk, niter=I, nysample=Yn, seed=seed, n_global=nx, x_offset=2).cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([a1, a2]), a)


@pytest.mark.parametrize("q,Sp,K,Yn,I", [(1, 2, 5, 1, 1), (1, 30, 37, 7, 2), (2, 3, 70, 5, 1), (3, 9, 20, 4, 2),
                                         (4, 17, 33, 9, 1), (8, 32, 43, 8, 1), (8, 40, 21, 3, 1), (5, 100, 12, 2, 1),
                                         (6, 8, 64, 16, 1)])
@pytest.mark.parametrize("use_seed", [False, True])
def test_tes_sp_whitened_kernel_matches_direct_kernel(monkeypatch, q, Sp, K, Yn, I, use_seed):
    """csrc/sp2.cuh (whitened residuals nu_j + z - nu_j', fixed LSE shift, table-driven exp; the default) against the
    direct kernel (y = mu + L z, triangular solve per log-density, max-shifted LSE, libm exp) on the same inputs:
    lane-group shapes G = 4...32, several rounds of components, ragged validity masks, ny not a multiple of the
    register block.  The direct kernel is the one the oracle tests pinned first; both are also checked against the
    oracle by the tests above (whichever is the default)."""
    from tes_b200 import ops
    rs = np.random.RandomState(1000 * q + Sp)
    d, n, T, nx = 3, 7, Sp, 5
    l, sigma, sigma0 = np.array([4.0, 2.5, 6.0]), 1.1, 1e-3
    X, Y, test_xs = rs.rand(n, d), rs.randn(n, 1) * 0.4, rs.rand(T, d)
    ls, sg, s0 = l.reshape(1, d), np.array([sigma]), np.array([sigma0])
    _, invpNK = onp.get_pNK_test_obs(ls, sg, s0, 1, X, test_xs)
    p = rs.rand(Sp) + 0.05
    p /= p.sum()
    P = rs.randn(Sp, T, K) * 0.7
    msk = rs.rand(Sp, K) < 0.7
    msk[:, 0] = True
    msk[0, :] = True
    xb = rs.rand(nx, q, d) if q > 1 else rs.rand(nx, d)
    kw = dict(niter=I, nysample=Yn)
    if use_seed:
        kw.update(seed=99, n_global=nx + 3, x_offset=2)
    else:
        kw.update(z=rs.normal(size=(I, Yn, Sp, nx, K) + ((q,) if q > 1 else ())))
    monkeypatch.setenv("TES_SP_METHOD", "1")
    a = ops.sp_acq(xb, test_xs, X, Y, l, sigma, sigma0, invpNK[0], p, P, msk, **kw).cpu().numpy()
    monkeypatch.setenv("TES_SP_METHOD", "2")
    b = ops.sp_acq(xb, test_xs, X, Y, l, sigma, sigma0, invpNK[0], p, P, msk, **kw).cpu().numpy()
    assert np.all(np.isfinite(a))
    np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-11)


def test_tes_sp_whitened_kernel_far_components_and_nan():
    """Components tens of standard deviations apart underflow in the fixed-shift LSE (-inf) where the max-shifted
    reference gives ~ -1e3: both contribute exp(.) = 0 to the outer LSE.  A NaN joint sample propagates to NaN."""
    from tes_b200 import ops
    pr = make_problem(seed=3, d=2, n=5, T=6, nx=4, hyp=BRANIN, mode="sample", nsample=200)
    p = pr["max_probs"][0]
    keep = p > 0
    P, msk = pr["v0"][0][keep].copy(), pr["v1"][0][keep]
    Sp, K = P.shape[0], P.shape[2]
    P[0] += 400.0                    # maximizer 0's joint samples sit ~1e4 sigma away from the others
    z = pr["rs"].normal(size=(1, 2, 3, Sp, 4, K))
    v0 = pr["v0"].copy()
    v0[0][np.where(keep)[0][0]] += 400.0
    acq = ops.sp_acq(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                     p[keep], P, msk, niter=2, nysample=3, z=z[0]).cpu().numpy()
    ref = onp.evaluate_sample_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, 4, 5, 1, 3,
                                 pr["test_xs"], pr["max_probs"], v0, pr["v1"], pr["invpNK"], niteration=2, z=z)
    _close(acq, ref)
    P[1, 0, 0] = np.nan
    acq = ops.sp_acq(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0],
                     p[keep], P, msk, niter=2, nysample=3, z=z[0]).cpu().numpy()
    assert np.all(np.isnan(acq))


@pytest.mark.parametrize("hyp,d,n,T,nx,mode", [(HART6, 6, 30, 20, 40_001, "empirical"), (BRANIN, 2, 6, 9, 19_000, "empirical"),
                                               (PH, 2, 20, 7, 400, "ep"), (HART6, 6, 40, 70, 20_000, "empirical")])
@pytest.mark.parametrize("tri", ["1", "0"])
def test_tes_ep_screen_bound_holds(hyp, d, n, T, nx, mode, tri, monkeypatch):
    """ops.ep_acq_screen: the 3xFP16 tcgen05 screen of the deterministic criterion stays within its own proven bound of
    the fp64 criterion on every candidate, and the bound is small enough to prune.  tri = 1: triangular GEMM with the
    tile of M as the A operand (default); 0: index-pair GEMM with the A operand generated on the SM."""
    from tes_b200 import ops
    monkeypatch.setenv("TES_QUAD_TRI", tri)
    pr = make_problem(seed=11 + T, d=d, n=n, T=T, nx=nx, hyp=hyp, mode=mode, nsample=max(600, 40 * T))
    p, pm, pc = _ep_inputs(pr)
    exact = ops.ep_acq(ops.EP_DETERMINISTIC, pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"],
                       pr["invpNK"][0], pr["invK"][0], p, pm, pc).cpu().numpy()
    acq, eps = ops.ep_acq_screen(pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"],
                                 pr["invpNK"][0], pr["invK"][0], p, pc)
    acq, eps = acq.cpu().numpy(), eps.cpu().numpy()
    fin = np.isfinite(eps) & np.isfinite(exact)
    assert fin.mean() > 0.99
    err = np.abs(acq - exact)[fin]
    assert np.all(err <= eps[fin]), (err.max(), eps[fin][np.argmax(err - eps[fin])])
    assert np.median(eps[fin]) < 1e-3 * max(1.0, np.abs(exact[fin]).max())
    # the screen is useful: the bound is far above the actual error but far below the spread of the criterion
    assert err.max() < 1e-4 * max(1.0, np.abs(exact[fin]).max())


def test_tes_ep_screen_more_than_128_test_points():
    """T > 128 test points (with S' <= 128 maximizer-conditioned posteriors): the triangular GEMM does not apply, the
    generator kernel runs; the bound still holds.  (The reference pipeline drops test points that are never the
    maximizer, so T = S' there; the operators accept the general case.)"""
    from tes_b200 import ops
    rs = np.random.RandomState(9)
    d, n, T, Sp, N = 6, 20, 140, 96, 5_000
    l, sigma, sigma0 = HART6["l"], HART6["sigma"], 1e-3
    X, Y, test_xs, x = rs.rand(n, d), rs.randn(n, 1) * 0.4, rs.rand(T, d), rs.rand(N, d)
    ls, sg, s0 = l.reshape(1, d), np.array([sigma]), np.array([sigma0])
    invK = onp.precomputeInvKs(ls, sg, s0, X)
    _, invpNK = onp.get_pNK_test_obs(ls, sg, s0, 1, X, test_xs)
    p = rs.rand(Sp)
    p /= p.sum()
    covs = []
    for _ in range(Sp):
        A = rs.randn(T, 2 * T)
        covs.append(A @ A.T / (2 * T) * 0.05)
    covs = np.stack(covs)
    exact = ops.ep_acq(ops.EP_DETERMINISTIC, x, test_xs, X, Y, l, sigma, sigma0, invpNK[0], invK[0], p,
                       np.zeros((Sp, T)), covs).cpu().numpy()
    acq, eps = [t.cpu().numpy() for t in ops.ep_acq_screen(x, test_xs, X, Y, l, sigma, sigma0, invpNK[0], invK[0], p, covs)]
    fin = np.isfinite(eps) & np.isfinite(exact)
    assert fin.mean() > 0.9
    assert np.all(np.abs(acq - exact)[fin] <= eps[fin])


def test_tes_ep_screen_fp32_bound_dominates_fp64_bound(monkeypatch):
    """The bound H = sum_a |M_a| sqrt(Sigma_aa) is formed on the fp32 pipe inside ep_det_screen_h_kernel (operands rounded
    up, result inflated by the fp32 summation error); TES_SCREEN_H=0 selects the fp64 GEMM.  eps(fp32) >= eps(fp64) on
    every candidate and within 2 % of it; the screen values agree to rounding."""
    from tes_b200 import ops
    pr = make_problem(seed=23, d=6, n=30, T=37, nx=9_001, hyp=HART6, mode="empirical", nsample=1500)
    p, pm, pc = _ep_inputs(pr)
    run = lambda: [t.cpu().numpy() for t in ops.ep_acq_screen(  # noqa: E731
        pr["x"], pr["test_xs"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invpNK"][0], pr["invK"][0], p, pc)]
    a32, e32 = run()
    monkeypatch.setenv("TES_SCREEN_H", "0")
    a64, e64 = run()
    np.testing.assert_allclose(a32, a64, rtol=1e-12, atol=1e-13)
    fin = np.isfinite(e64)
    assert np.array_equal(fin, np.isfinite(e32))
    assert np.all(e32[fin] >= e64[fin] * (1 - 1e-12))
    # ... and not much looser: the fused kernel also replaces 1/2 log(w / (w - delta)) by 1/2 delta / (w - delta) and adds
    # 1e-13 of the summed |terms| instead of 1e-13 |acq|
    small = fin & (e64 < 1e-2)
    assert small.mean() > 0.9
    assert np.all(e32[small] <= e64[small] * 1.02 + 1e-11)


def test_tes_step_with_screen_selects_the_same_query():
    """acquisition.tes_step with the screened criterion (candidates that can still be the arg-max re-evaluated in fp64)
    returns the same query index, point and VALUE as the fp64 criterion on all candidates."""
    from tes_b200 import acquisition
    rs = np.random.RandomState(5)
    d, n, S, F, N = 6, 24, 48, 128, 70_001
    l, sigma, sigma0 = HART6["l"], HART6["sigma"], 1e-3
    cand, X = rs.rand(N, d), rs.rand(n, d)
    Y = rs.randn(n, 1) * 0.5
    theta, W, b = onp.draw_random_init_weights_features(d, S, F, X, Y, l.reshape(1, d), sigma, sigma0, rs.randn(S, F, d),
                                                        rs.rand(S, F, 1), rs.randn(S, F, 1))
    kw = dict(criterion="ftl", mode="empirical", deterministic=True, ntestsample=3000, seed=3, t_max=40)
    a = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, screen=False, **kw)
    b_ = acquisition.tes_step(cand, X, Y, l, sigma, sigma0, theta, W, b, screen=True, **kw)
    assert a["query_idx"] == b_["query_idx"] and a["query_val"] == b_["query_val"]
    np.testing.assert_array_equal(a["query_x"], b_["query_x"])
    assert b_["screen_kept"] is not None and 1 <= b_["screen_kept"] <= N // 8
