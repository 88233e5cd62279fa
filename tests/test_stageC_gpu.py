"""Stage C on the device (csrc/maxstats.cu) against the numpy oracle and the reference's own outputs
(tests/golden/maximizer_stats.npz).  Calls go through the C ABI (tes_b200.ops / utils)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import tes_oracle_np as onp

pytestmark = pytest.mark.gpu


def test_normal_fill_matches_the_c_oracle():
    from tes_b200 import ops
    for first, count in [(0, 1001), (7, 4096), (12345, 2)]:
        got = ops.normal_fill(99, 3, ops.STREAM_JOINT_Z, count, first_elem=first).cpu().numpy()
        ref = c_oracle.normals(99, 3, 5, first, count)
        # log/sin/cos come from the platform libm on each side (tes_spec.h): ~1 ulp, not bit-exact
        np.testing.assert_allclose(got, ref, rtol=0, atol=5e-15)
    z = ops.normal_fill(1, 0, ops.STREAM_JOINT_Z, 1 << 20).cpu().numpy()
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3


@pytest.mark.parametrize("d", [1, 2, 6, 8, 10, 19])
def test_dedup_is_identical_to_the_reference_walk(d):
    from tes_b200 import acquisition, ops
    rs = np.random.RandomState(d)
    centers = rs.rand(40, d)
    xs = centers[rs.randint(0, 40, size=700)] + rs.randn(700, d) * 4e-4   # many pairs near the 1e-3 threshold
    mask, leader = ops.dedup(xs, 1e-3)
    ref_mask = onp.get_duplicate_mask(xs, 1e-3)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask.astype(np.int32))
    np.testing.assert_array_equal(acquisition.select_test_points(xs, 1e-3), onp.remove_duplicates(xs, 1e-3))
    lead = leader.cpu().numpy()
    assert np.all(lead[ref_mask == 0] == np.where(ref_mask == 0)[0]) and np.all(lead[ref_mask == 1] < np.where(ref_mask == 1)[0])


def test_dedup_golden(golden):
    from tes_b200 import acquisition
    g = golden("maximizer_stats")
    np.testing.assert_array_equal(acquisition.select_test_points(g["dedup_in"].copy(), 1e-3), g["dedup_out"])


def test_reference_named_numpy_helpers_run_on_the_device(golden):
    """remove_duplicates_np / get_duplicate_mask_np (utils.py:1554-1581), get_empirical_stat_from_samples
    (empirical_approximation.py:86-107) and sample_multivariate_normal_maxidx_np (utils.py:1654-1729) of the package
    are device-backed; pinned on the reference's own outputs."""
    from tes_b200 import _lib, utils
    from tes_b200 import empirical_approximation as ea
    g = golden("maximizer_stats")
    n0 = _lib.lib().tes_launch_count()
    np.testing.assert_array_equal(utils.remove_duplicates_np(g["dedup_in"].copy(), 1e-3), g["dedup_out"])
    np.testing.assert_array_equal(utils.get_duplicate_mask_np(g["dedup_in"].copy(), 1e-3),
                                  onp.get_duplicate_mask(g["dedup_in"].copy(), 1e-3))
    m, c = ea.get_empirical_stat_from_samples(g["emp_samples"], g["emp_weight"])
    np.testing.assert_allclose(m, g["emp_mean"], rtol=1e-13)
    np.testing.assert_allclose(c, g["emp_cov"], rtol=1e-12, atol=1e-15)
    m2, c2 = ea.get_empirical_stat_from_samples(g["emp_samples"])
    rm, rc = onp.get_empirical_stat_from_samples(g["emp_samples"])
    np.testing.assert_allclose(m2, rm, rtol=1e-13)
    np.testing.assert_allclose(c2, rc, rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError):
        ea.get_empirical_stat_from_samples(g["emp_samples"], 0.5 * np.ones(g["emp_samples"].shape[0]))
    for tag in ("t6", "t12"):
        mean, cov, z = g[tag + "_mean"][0], g[tag + "_cov"][0], g[tag + "_z"][0]
        A = onp.mvn_transform_matrix(cov)
        for rc_ in (True, False):
            ref = onp.sample_multivariate_normal_maxidx(mean.copy(), cov.copy(), z.shape[0], 2, z=z,
                                                        remove_correlated_dims=rc_, transform=A)
            got = utils.sample_multivariate_normal_maxidx_np(mean.copy(), cov.copy(), z.shape[0], 2, z=z,
                                                             remove_correlated_dims=rc_, transform=A)
            assert got[0] == ref[0] and got[2] == ref[2]
            np.testing.assert_array_equal(got[1], ref[1])
            np.testing.assert_allclose(got[3], ref[3], rtol=1e-12, atol=1e-13)
            np.testing.assert_array_equal(got[4], ref[4])
            np.testing.assert_array_equal(got[5], ref[5])
            u, sz = utils.sample_multivariate_normal_maxidx_np(mean.copy(), cov.copy(), z.shape[0], 2, get_sample=False,
                                                               z=z, remove_correlated_dims=rc_, transform=A)
            np.testing.assert_array_equal(u, ref[1])
            np.testing.assert_array_equal(sz, ref[5])
    assert _lib.lib().tes_launch_count() > n0


def test_select_test_points_t_max_matches_oracle_pipeline():
    from oracle import pipeline_np
    from tes_b200 import acquisition
    rs = np.random.RandomState(5)
    centers = rs.rand(30, 3)
    xs = centers[rs.randint(0, 30, size=400)]
    for t_max in (None, 5, 12, 64):
        np.testing.assert_array_equal(acquisition.select_test_points(xs, 1e-3, t_max),
                                      pipeline_np.select_test_points(xs, 1e-3, t_max))


@pytest.mark.parametrize("n,batch", [(1, 3), (2, 5), (7, 4), (64, 9), (125, 2), (160, 1), (200, 2)])
def test_sym_eig(n, batch):
    from tes_b200 import ops
    rs = np.random.RandomState(n)
    Z = rs.randn(batch, n, max(n // 2, 1) if n > 100 else 3 * n)   # n > 100: rank-deficient PSD (zero eigenvalues)
    A = Z @ np.transpose(Z, (0, 2, 1)) + (0.0 if n > 100 else 1e-4) * np.eye(n)
    e, v, sweeps = ops.sym_eig(A, return_sweeps=True)
    e, v = e.cpu().numpy(), v.cpu().numpy()
    assert int(sweeps.max()) < 40
    scale = np.abs(np.linalg.eigvalsh(A)).max(axis=1)
    for b in range(batch):
        np.testing.assert_allclose(np.sort(e[b]), np.linalg.eigvalsh(A[b]), rtol=0, atol=1e-12 * scale[b])
        rec = (v[b] * e[b]) @ v[b].T
        np.testing.assert_allclose(rec, A[b], rtol=0, atol=2e-12 * scale[b])
        keep = np.abs(e[b]) > 1e-9 * scale[b]                          # null-space columns need not be normalised
        vv = v[b][:, keep]
        np.testing.assert_allclose(vv.T @ vv, np.eye(keep.sum()), rtol=0, atol=1e-11)


def test_color_factor_reproduces_the_covariance():
    from tes_b200 import ops
    rs = np.random.RandomState(0)
    X = rs.rand(40, 3)
    cov = onp.computeKmm(X, np.array([30.0, 20.0, 10.0]), 1.3) if hasattr(onp, "computeKmm") else None
    if cov is None:
        Z = rs.randn(40, 60)
        cov = Z @ Z.T
    A = ops.color_factor(cov).cpu().numpy()
    np.testing.assert_allclose(A @ A.T, cov, rtol=0, atol=1e-12 * np.abs(cov).max())


@pytest.mark.parametrize("tag", ["t6", "t12"])
@pytest.mark.parametrize("mode", ["sample", "empirical", "ep"])
def test_device_get_testidxs_stats_matches_the_reference_outputs(golden, tag, mode):
    """Same z and the same colouring factor (np.linalg.eig of the same matrix, as the reference computes it) ->
    the reference's own outputs."""
    from tes_b200 import utils
    g = golden("maximizer_stats")
    cov = g[tag + "_cov"].copy()
    A = onp.mvn_transform_matrix(cov[0])
    out = utils.get_testidxs_stats(1, g[tag + "_test_xs"].copy(), g[tag + "_mean"].copy(), cov, mode=mode,
                                   nsample=int(g[tag + "_nsample"]), n_min_sample=2, z=g[tag + "_z"], transforms=[A])
    if mode == "ep":
        np.testing.assert_allclose(out[4][0], g[tag + "_ep_mean"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out[5][0], g[tag + "_ep_cov"], rtol=1e-9, atol=1e-11)
        return
    for i, o in enumerate(out):
        ref = g["%s_%s_out%d" % (tag, mode, i)]
        assert isinstance(o, np.ndarray) and o.shape == ref.shape
        if ref.dtype == bool:
            np.testing.assert_array_equal(o, ref)
        else:
            np.testing.assert_allclose(o, ref, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("mode", ["sample", "empirical"])
@pytest.mark.parametrize("T,nsample,nhyp", [(5, 200, 1), (33, 3000, 1), (128, 8192, 1), (9, 500, 3)])
def test_device_get_testidxs_stats_vs_oracle(mode, T, nsample, nhyp):
    """Random GP posteriors with near-duplicate test points (exercises the correlated-dimension mask), several
    hyper-parameter samples (exercises the pruning across them)."""
    from tes_b200 import utils
    rs = np.random.RandomState(T)
    d = 3
    test_xs = rs.rand(T, d)
    if T > 6:
        test_xs[5] = test_xs[2] + 1e-7            # correlation ~ 1: dimension 5 is suppressed
    X, Y = rs.rand(12, d), rs.randn(12, 1)
    means, covs, trs = [], [], []
    for i in range(nhyp):
        l = np.array([20.0, 9.0, 14.0]) * (1.0 + 0.3 * i)
        ls, sg, s0 = l.reshape(1, d), np.array([1.1]), np.array([1e-3])
        invK = onp.precomputeInvKs(ls, sg, s0, X)
        m, c = onp.compute_mean_var_f(test_xs, X, Y, l, 1.1, 1e-3, invK[0], fullcov=True)
        means.append(m.reshape(T))
        covs.append(c)
        trs.append(onp.mvn_transform_matrix(c))
    means, covs = np.stack(means), np.stack(covs)
    z = rs.normal(size=(nhyp, nsample, T, 1))
    ref = onp.get_testidxs_stats(nhyp, test_xs.copy(), means.copy(), covs.copy(), mode=mode, nsample=nsample,
                                 n_min_sample=2, z=z, transforms=trs)
    out = utils.get_testidxs_stats(nhyp, test_xs.copy(), means.copy(), covs.copy(), mode=mode, nsample=nsample,
                                   n_min_sample=2, z=z, transforms=trs)
    assert len(ref) == len(out)
    for o, r in zip(out, ref):
        r = np.asarray(r)
        assert o.shape == r.shape
        if r.dtype == bool:
            np.testing.assert_array_equal(o, r)
        else:
            np.testing.assert_allclose(o, r, rtol=1e-9, atol=1e-11)
    # CUDA tensors in -> CUDA tensors out, same values
    out_t = utils.get_testidxs_stats(nhyp, torch.from_numpy(test_xs).cuda(), torch.from_numpy(means).cuda(),
                                     torch.from_numpy(covs).cuda(), mode=mode, nsample=nsample, n_min_sample=2, z=z,
                                     transforms=trs)
    for o, t in zip(out, out_t):
        assert t.is_cuda
        np.testing.assert_array_equal(o, t.cpu().numpy())


def test_device_joint_samples_with_own_colouring_factor_have_the_right_law():
    """Without `transforms` the factor comes from the device eigen-solver: same distribution, not the same draws
    as LAPACK's eig; check the bucket probabilities against the host path statistically."""
    from tes_b200 import utils
    rs = np.random.RandomState(1)
    T, d = 8, 2
    test_xs, X, Y = rs.rand(T, d), rs.rand(6, d), rs.randn(6, 1)
    l = np.array([12.0, 7.0])
    invK = onp.precomputeInvKs(l.reshape(1, d), np.array([1.0]), np.array([1e-3]), X)
    m, c = onp.compute_mean_var_f(test_xs, X, Y, l, 1.0, 1e-3, invK[0], fullcov=True)
    a = utils.get_testidxs_stats(1, test_xs, m.reshape(1, T), c[None], mode="empirical", nsample=200000, seed=4)
    b = onp.get_testidxs_stats(1, test_xs, m.reshape(1, T), c[None], mode="empirical", nsample=200000,
                               z=rs.normal(size=(1, 200000, T, 1)))
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_allclose(a[1], b[1], atol=6e-3)
    np.testing.assert_allclose(a[4], b[4], atol=3e-2)


# ---- RFF posterior weight draw on the device (csrc/rffdraw.cu, optfunc.py:303-385) ------------------------------
@pytest.mark.parametrize("tag,xdim", [("nltF", 2), ("ngeF", 2), ("d6", 6)])
def test_device_rff_weight_draw_matches_the_reference_outputs(golden, tag, xdim):
    from tes_b200 import optfunc
    g = golden("rff")
    S, F = g[tag + "_theta"].shape[:2]
    theta, W, b = optfunc.draw_random_init_weights_features_np(
        xdim, S, F, g[tag + "_X"], g[tag + "_Y"], g[tag + "_l"], float(g[tag + "_sigma"]), float(g[tag + "_sigma0"]),
        draws=(g[tag + "_randn_W"], g[tag + "_rand_b"], g[tag + "_randn_noise"]))
    np.testing.assert_array_equal(W, g[tag + "_W"])        # elementwise products: bit-exact
    np.testing.assert_array_equal(b, g[tag + "_b"])
    # theta goes through Sigma^-1 (cond ~ trace/sigma0): 1e-8 relative to the largest weight
    np.testing.assert_allclose(theta, g[tag + "_theta"], rtol=0, atol=1e-8 * np.abs(g[tag + "_theta"]).max())


@pytest.mark.parametrize("S,F,n,d,sigma0", [(16, 256, 64, 6, 1e-4), (5, 300, 5, 2, 1e-4), (3, 512, 200, 4, 1e-2),
                                            (4, 48, 60, 3, 1e-2), (2, 64, 64, 2, 1e-3)])
def test_device_rff_weight_draw_vs_oracle(S, F, n, d, sigma0):
    from tes_b200 import optfunc
    rs = np.random.RandomState(S * 1000 + n)
    X, Y = rs.rand(n, d), rs.randn(n, 1)
    l = 1.0 + 8.0 * rs.rand(d)
    draws = (rs.randn(S, F, d), rs.rand(S, F, 1), rs.randn(S, F, 1))
    ref = onp.draw_random_init_weights_features(d, S, F, X, Y, l.reshape(1, d), 1.3, sigma0, *draws)
    got = optfunc.draw_random_init_weights_features_np(d, S, F, X, Y, l.reshape(1, d), 1.3, sigma0, draws=draws)
    np.testing.assert_array_equal(got[1], ref[1])
    np.testing.assert_array_equal(got[2], ref[2])
    np.testing.assert_allclose(got[0], ref[0], rtol=0, atol=1e-7 * np.abs(ref[0]).max())


def test_device_rff_weight_draw_from_seed_is_a_posterior_sample():
    """Counter-based draws on the device: the function samples must interpolate the observations (posterior with
    small noise) -- a property test at a size the host path would take seconds for."""
    from tes_b200 import ops, optfunc
    rs = np.random.RandomState(0)
    n, d, S, F = 40, 6, 256, 2048
    X = rs.rand(n, d)
    l = np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949])
    Y = np.sin(3.0 * X.sum(axis=1, keepdims=True))
    theta, W, b = optfunc.draw_random_init_weights_features(d, S, F, X, Y, l.reshape(1, d), 1.423, 1e-4, seed=12)
    assert theta.is_cuda and theta.shape == (S, F, 1)
    f = ops.rff_eval(X, W, b, theta, 1.423).cpu().numpy()           # (S, n)
    assert np.abs(f - Y.reshape(1, n)).max() < 0.2                    # sd of the posterior at the data ~ sqrt(sigma0)
    u = ops.uniform_fill(12, 0, ops.STREAM_RFF_B, 1 << 16).cpu().numpy()
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    raw = c_oracle.philox_raw(12, 0, 7, 5)                             # elements 10, 11 of the uniform stream
    k0 = ((raw[1] << 32) | raw[0]) >> 11
    assert u[10] == (k0 + 0.5) * 2.0 ** -53


# ---- batched EP on the device (csrc/ep.cu, ep.py:73-204) --------------------------------------------------------
@pytest.mark.parametrize("tag", ["t6", "t12"])
def test_device_ep_matches_the_reference_outputs(golden, tag):
    from tes_b200 import ep
    g = golden("maximizer_stats")
    mean_k, cov_k = g[tag + "_sample_out2"], g[tag + "_sample_out3"]
    pm, pc = ep.approximate_EP_all(mean_k[0], cov_k[0], 1e-9, 100)
    np.testing.assert_allclose(pm, g[tag + "_ep_mean"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(pc, g[tag + "_ep_cov"], rtol=1e-8, atol=1e-10)
    m, c = ep.approximate_EP_np(1, mean_k[0].reshape(-1, 1), cov_k[0], 1e-9, 100)     # the reference's signature
    np.testing.assert_allclose(m.squeeze(), g[tag + "_ep_mean"][1], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(c, g[tag + "_ep_cov"][1], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("T,nhyp,max_niter,res", [(2, 1, 100, 1e-9), (3, 2, 100, 1e-9), (9, 1, 1000, 1e-14),
                                                  (24, 2, 100, 1e-9), (61, 1, 60, 1e-9), (170, 1, 6, 1e-9)])
def test_device_ep_vs_oracle(T, nhyp, max_niter, res):
    """Random GP posteriors at well-separated test points (cond(Sigma) moderate so that the two inverses agree to
    1e-9); T = 170 exercises the global-memory matrix path; same iteration counts as the oracle."""
    from tes_b200 import ops
    rs = np.random.RandomState(T)
    d = 2
    means, covs = [], []
    for i in range(nhyp):
        xs = rs.rand(T, d) * (4.0 if T < 100 else 30.0)
        X, Y = rs.rand(5, d) * 4.0, rs.randn(5, 1)
        l = np.array([3.0, 5.0]) * (1 + i)
        invK = onp.precomputeInvKs(l.reshape(1, d), np.array([1.0]), np.array([1e-2]), X)
        m, c = onp.compute_mean_var_f(xs, X, Y, l, 1.0, 1e-2, invK[0], fullcov=True)
        means.append(m.reshape(T))
        covs.append(c + 1e-6 * np.eye(T))
    means, covs = np.stack(means), np.stack(covs)
    pm, pc, nit = ops.ep_moments(means, covs, res, max_niter, return_niter=True)
    pm, pc = pm.cpu().numpy(), pc.cpu().numpy()
    for i in range(nhyp):
        for j in (range(T) if T <= 24 else [0, 1, T // 2, T - 1]):
            m, c = onp.approximate_EP(j, means[i].reshape(-1, 1), covs[i], res, max_niter)
            scale = np.abs(covs[i]).max()
            np.testing.assert_allclose(pm[i, j], m.squeeze(), rtol=0, atol=2e-8 * max(1.0, np.abs(m).max()))
            np.testing.assert_allclose(pc[i, j], c, rtol=0, atol=2e-8 * scale)
    assert int(nit.min()) >= 1


def test_device_ep_kernels_agree(monkeypatch):
    """T <= 32 has three engines (csrc/ep.cu): four warps per maximizer (default), one warp, one CTA; same pivot rule
    and element arithmetic, so the moments agree to rounding and the iteration counts are identical."""
    from tes_b200 import ops
    rs = np.random.RandomState(30)
    T, d = 30, 2
    xs = rs.rand(T, d) * 4.0
    X, Y = rs.rand(6, d) * 4.0, rs.randn(6, 1)
    l = np.array([3.0, 5.0])
    invK = onp.precomputeInvKs(l.reshape(1, d), np.array([1.0]), np.array([1e-2]), X)
    m, c = onp.compute_mean_var_f(xs, X, Y, l, 1.0, 1e-2, invK[0], fullcov=True)
    means, covs = m.reshape(1, T), (c + 1e-6 * np.eye(T))[None]
    res = {}
    for mode in ("0", "3", "1"):
        monkeypatch.setenv("TES_EP_MODE", mode)
        pm, pc, nit = ops.ep_moments(means, covs, 1e-9, 100, return_niter=True)
        res[mode] = (pm.cpu().numpy(), pc.cpu().numpy(), nit.cpu().numpy())
    for mode in ("3", "1"):
        np.testing.assert_array_equal(res[mode][2], res["0"][2])
        np.testing.assert_allclose(res[mode][0], res["0"][0], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(res[mode][1], res["0"][1], rtol=1e-10, atol=1e-12)
    j = 7
    om, oc = onp.approximate_EP(j, means[0].reshape(-1, 1), covs[0], 1e-9, 100)
    np.testing.assert_allclose(res["0"][0][0, j], om.squeeze(), rtol=0, atol=2e-8)
    np.testing.assert_allclose(res["0"][1][0, j], oc, rtol=0, atol=2e-8 * np.abs(covs).max())


def test_device_ep_mode_through_get_testidxs_stats_cuda_tensors(golden):
    from tes_b200 import utils
    g = golden("maximizer_stats")
    tag = "t12"
    cov = g[tag + "_cov"].copy()
    out = utils.get_testidxs_stats(1, torch.from_numpy(g[tag + "_test_xs"]).cuda(),
                                   torch.from_numpy(g[tag + "_mean"]).cuda(), torch.from_numpy(cov).cuda(), mode="ep",
                                   nsample=int(g[tag + "_nsample"]), n_min_sample=2, z=g[tag + "_z"],
                                   transforms=[onp.mvn_transform_matrix(cov[0])])
    assert all(o.is_cuda for o in out)
    np.testing.assert_allclose(out[4][0].cpu().numpy(), g[tag + "_ep_mean"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(out[5][0].cpu().numpy(), g[tag + "_ep_cov"], rtol=1e-8, atol=1e-10)


def test_weight_draw_of_a_slice_of_samples_is_the_slice_of_the_draw():
    """Multi-GPU steps shard the weight draw by function sample (acquisition.tes_step): drawing samples [lo, hi) alone gives
    the bits of the full draw's rows [lo, hi) -- every kernel of stage A0 treats the samples independently."""
    from tes_b200 import optfunc
    rs = np.random.RandomState(8)
    S, F, n, d = 24, 96, 20, 6
    X, Y = rs.rand(n, d), rs.randn(n, 1) * 0.4
    l, sigma, sigma0 = np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]), 1.423, 1e-4
    draws = tuple(torch.from_numpy(a).cuda() for a in (rs.randn(S, F, d), rs.rand(S, F, 1), rs.randn(S, F, 1)))
    full = optfunc.draw_random_init_weights_features(d, S, F, X, Y, l, sigma, sigma0, draws=draws)
    for lo, hi in ((0, 8), (8, 16), (16, 24), (5, 6)):
        part = optfunc.draw_random_init_weights_features(d, hi - lo, F, X, Y, l, sigma, sigma0,
                                                         draws=tuple(t[lo:hi].contiguous() for t in draws))
        for a, b in zip(full, part):
            assert torch.equal(a[lo:hi], b)
