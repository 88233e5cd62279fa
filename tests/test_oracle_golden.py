"""Pin the numpy oracle (oracle/tes_oracle_np.py) against the reference's own outputs.

Fixtures in tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
reference's numpy functions unmodified from /root/reference (build container only).
"""
import numpy as np
import pytest

from oracle import tes_oracle_np as onp


# ---- the reference's only shipped known-answer data: objective maxima --------------------------
@pytest.mark.parametrize("name,tol", [("pH", 5e-9), ("log10P", 5e-9), ("log10K", 5e-7)])
def test_brooms_barn_known_answers(golden, name, tol):
    g = golden("kat_objectives")
    X, Y, hyp = g[name + "_X"], g[name + "_Y"], g[name + "_hyp"]
    sigma, l, sigma0 = hyp[0], hyp[1:3], hyp[3]
    f = onp.compute_mean_f_np(g[name + "_maximizer"], X, Y, l, sigma, sigma0)
    # functions.py:397-398 (pH), :363-364 (log10P), :329-330 (log10K, shipped to 4e-7 abs)
    assert abs(f[0] - float(g[name + "_maximum"])) < tol
    np.testing.assert_allclose(f, g[name + "_f_at_maximizer"], rtol=0, atol=1e-12)
    fx = onp.compute_mean_f_np(g[name + "_xs"], X, Y, l, sigma, sigma0)
    np.testing.assert_allclose(fx, g[name + "_f_on_xs"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_se_ard_matches_reference_numpy(golden, tag):
    g = golden("se_ard")
    A, B, l, sigma = g[tag + "_A"], g[tag + "_B"], g[tag + "_l"], float(g[tag + "_sigma"])
    np.testing.assert_array_equal(onp.computeKnm(A, B, l, sigma), g[tag + "_Knm"])
    np.testing.assert_array_equal(onp.computeKmm(A, l, sigma), g[tag + "_Kmm"])


@pytest.mark.parametrize("tag,xdim", [("nltF", 2), ("ngeF", 2), ("d6", 6)])
def test_rff_draw_and_function_samples(golden, tag, xdim):
    g = golden("rff")
    theta_ref, W_ref, b_ref = g[tag + "_theta"], g[tag + "_W"], g[tag + "_b"]
    S, F = theta_ref.shape[0], theta_ref.shape[1]
    theta, W, b = onp.draw_random_init_weights_features(
        xdim, S, F, g[tag + "_X"], g[tag + "_Y"], g[tag + "_l"], float(g[tag + "_sigma"]),
        float(g[tag + "_sigma0"]), g[tag + "_randn_W"], g[tag + "_rand_b"], g[tag + "_randn_noise"])
    np.testing.assert_array_equal(W, W_ref)
    np.testing.assert_array_equal(b, b_ref)
    np.testing.assert_allclose(theta, theta_ref, rtol=1e-9, atol=1e-10)
    f = onp.rff_fvals(g[tag + "_xs"], theta_ref, W_ref, b_ref, float(g[tag + "_sigma"]))
    np.testing.assert_allclose(f, g[tag + "_fvals"], rtol=0, atol=1e-12)
    idx, val = onp.rff_topk(g[tag + "_xs"], theta_ref, W_ref, b_ref, float(g[tag + "_sigma"]), 3)
    np.testing.assert_array_equal(idx[:, 0], g[tag + "_argmax"])
    assert np.all(val[:, 0] >= val[:, 1]) and np.all(val[:, 1] >= val[:, 2])


@pytest.mark.parametrize("tag", ["t6", "t12"])
@pytest.mark.parametrize("mode", ["sample", "empirical"])
def test_get_testidxs_stats(golden, tag, mode):
    g = golden("maximizer_stats")
    out = onp.get_testidxs_stats(
        1, g[tag + "_test_xs"].copy(), g[tag + "_mean"].copy(), g[tag + "_cov"].copy(), mode=mode,
        nsample=int(g[tag + "_nsample"]), n_min_sample=2, z=g[tag + "_z"],
        transforms=g[tag + "_transform"][None])
    for i, o in enumerate(out):
        ref = g["%s_%s_out%d" % (tag, mode, i)]
        assert np.asarray(o).shape == ref.shape, (i, np.asarray(o).shape, ref.shape)
        if ref.dtype == bool:
            np.testing.assert_array_equal(o, ref)
        else:
            np.testing.assert_allclose(o, ref, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag", ["t6", "t12"])
def test_transform_matrix_reproduces(golden, tag):
    # np.linalg.eig is LAPACK dgeev: same library, same matrix -> same factor in this container
    g = golden("maximizer_stats")
    A = onp.mvn_transform_matrix(g[tag + "_cov"][0])
    np.testing.assert_allclose(A @ A.T, g[tag + "_transform"] @ g[tag + "_transform"].T,
                               rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", ["t6", "t12"])
def test_ep_matches_reference(golden, tag):
    g = golden("maximizer_stats")
    mean_k, cov_k = g[tag + "_sample_out2"], g[tag + "_sample_out3"]
    T = mean_k.shape[1]
    for j in range(T):
        m, c = onp.approximate_EP(j, mean_k[0].reshape(-1, 1), cov_k[0], 1e-9, 100)
        np.testing.assert_allclose(m.squeeze(), g[tag + "_ep_mean"][j], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(c, g[tag + "_ep_cov"][j], rtol=1e-9, atol=1e-11)


def test_empirical_and_dedup(golden):
    g = golden("maximizer_stats")
    m, c = onp.get_empirical_stat_from_samples(g["emp_samples"], g["emp_weight"])
    np.testing.assert_array_equal(m, g["emp_mean"])
    np.testing.assert_array_equal(c, g["emp_cov"])
    np.testing.assert_array_equal(onp.remove_duplicates(g["dedup_in"].copy(), 1e-3), g["dedup_out"])
