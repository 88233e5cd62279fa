"""Host-side logic of the product package (the numpy halves the reference also runs on the host)
against the reference's own outputs (golden fixtures).  No GPU needed."""
import numpy as np
import pytest


@pytest.mark.parametrize("tag", ["t6", "t12"])
@pytest.mark.parametrize("mode", ["sample", "empirical"])
def test_get_testidxs_stats_matches_reference(golden, tag, mode):
    from tes_b200 import utils
    g = golden("maximizer_stats")
    out = utils.get_testidxs_stats_host(1, g[tag + "_test_xs"].copy(), g[tag + "_mean"].copy(), g[tag + "_cov"].copy(),
                                   mode=mode, nsample=int(g[tag + "_nsample"]), n_min_sample=2, z=g[tag + "_z"])
    for i, o in enumerate(out):
        ref = g["%s_%s_out%d" % (tag, mode, i)]
        assert np.asarray(o).shape == ref.shape
        if ref.dtype == bool:
            np.testing.assert_array_equal(o, ref)
        else:
            np.testing.assert_allclose(o, ref, rtol=1e-11, atol=1e-12)


def test_get_testidxs_stats_uses_global_rng_like_the_reference(golden):
    from tes_b200 import utils
    g = golden("maximizer_stats")
    np.random.seed(42)  # the fixture was generated with np.random.seed(42) before the reference call
    out = utils.get_testidxs_stats_host(1, g["t6_test_xs"].copy(), g["t6_mean"].copy(), g["t6_cov"].copy(),
                                   mode="sample", nsample=int(g["t6_nsample"]), n_min_sample=2)
    np.testing.assert_allclose(out[4], g["t6_sample_out4"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("tag", ["t6", "t12"])
def test_ep_mirror(golden, tag):
    from tes_b200 import ep, utils
    g = golden("maximizer_stats")
    mean_k, cov_k = g[tag + "_sample_out2"], g[tag + "_sample_out3"]
    for j in range(mean_k.shape[1]):
        m, c = ep.approximate_EP_host(j, mean_k[0].reshape(-1, 1), cov_k[0], 1e-9, 100)
        np.testing.assert_allclose(m.squeeze(), g[tag + "_ep_mean"][j], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(c, g[tag + "_ep_cov"][j], rtol=1e-9, atol=1e-11)
    # mode='ep' end to end (the call utils.py:1820-1825 intends)
    out = utils.get_testidxs_stats_host(1, g[tag + "_test_xs"].copy(), g[tag + "_mean"].copy(), g[tag + "_cov"].copy(),
                                   mode="ep", nsample=int(g[tag + "_nsample"]), n_min_sample=2, z=g[tag + "_z"])
    np.testing.assert_allclose(out[4][0], g[tag + "_ep_mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(out[5][0], g[tag + "_ep_cov"], rtol=1e-9, atol=1e-11)


def test_empirical_and_dedup_mirror(golden):
    from tes_b200 import empirical_approximation as ea
    from tes_b200 import utils
    g = golden("maximizer_stats")
    m, c = ea.get_empirical_stat_from_samples(g["emp_samples"], g["emp_weight"])
    np.testing.assert_allclose(m, g["emp_mean"], rtol=1e-14)
    np.testing.assert_allclose(c, g["emp_cov"], rtol=1e-13, atol=1e-16)
    np.testing.assert_array_equal(utils.remove_duplicates_np(g["dedup_in"].copy(), 1e-3), g["dedup_out"])


@pytest.mark.parametrize("tag,xdim", [("nltF", 2), ("ngeF", 2), ("d6", 6)])
def test_rff_weight_draw_mirror(golden, tag, xdim):
    from tes_b200 import optfunc
    g = golden("rff")
    S, F = g[tag + "_theta"].shape[:2]
    theta, W, b = optfunc.draw_random_init_weights_features_host(
        xdim, S, F, g[tag + "_X"], g[tag + "_Y"], g[tag + "_l"], float(g[tag + "_sigma"]), float(g[tag + "_sigma0"]),
        draws=(g[tag + "_randn_W"], g[tag + "_rand_b"], g[tag + "_randn_noise"]))
    np.testing.assert_array_equal(W, g[tag + "_W"])
    np.testing.assert_array_equal(b, g[tag + "_b"])
    np.testing.assert_allclose(theta, g[tag + "_theta"], rtol=1e-9, atol=1e-10)
