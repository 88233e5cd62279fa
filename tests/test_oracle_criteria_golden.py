"""Pin the numpy oracle's five TES criteria and TF-half helpers on outputs of the REFERENCE's own
criteria/*.py + utils.py, run unmodified on the numpy TF-1.14/TFP-0.7 shim (oracle/tf_shim) by
tests/golden/make_golden_criteria.py (build container).  Fixture: tests/golden/criteria.npz.

Tolerances: the oracle restates the same fp64 arithmetic with different association in a few
matmuls (einsum vs tiled batch matmul), so values agree to ~1e-12; 1e-9 relative is asserted.
Batch TES_ep is ~0 by construction: absolute 1e-12.
"""
import numpy as np
import pytest

from oracle import tes_oracle_np as onp

CASES = ["c1", "c2", "h2", "d6"]
RT, AT = 1e-9, 1e-12


class Case:
    def __init__(self, g, tag):
        self.g, self.tag = g, tag
        G = self.__getitem__
        self.ls, self.sigmas, self.sigma0s = G("ls"), G("sigmas"), G("sigma0s")
        self.X, self.Y, self.x, self.xq = G("X"), G("Y"), G("x"), G("xq")
        self.nhyp, self.d = self.ls.shape
        self.n, self.nx = self.X.shape[0], self.x.shape[0]
        self.I, self.Yn = int(G("I")), int(G("Y_n"))
        (self.test_xs, self.max_probs, self.mean_k, self.cov_k, self.samples,
         self.masks) = [G("sample_out%d" % k) for k in range(6)]
        self.emp_mean, self.emp_cov = G("empirical_out4"), G("empirical_out5")
        self.invK, self.invpNK = G("invK"), G("invpNK")

    def __getitem__(self, k):
        return self.g["%s_%s" % (self.tag, k)]

    def moments(self, which):
        return (self.emp_mean, self.emp_cov) if which == "emp" else (self["ep_mean"], self["ep_cov"])


@pytest.fixture(scope="module")
def case(golden):
    g = golden("criteria")
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = Case(g, tag)
        return cache[tag]
    return get


@pytest.mark.parametrize("tag", CASES)
def test_stage_b_helpers(case, tag):
    """utils.precomputeInvKs / compute_mean_var_f (TF halves, utils.py:786-815, 1838-1856) and
    get_pNK_test_obs (evaluate_mp.py:12-53)."""
    c = case(tag)
    invK = onp.precomputeInvKs(c.ls, c.sigmas, c.sigma0s, c.X)
    np.testing.assert_allclose(invK, c.invK, rtol=1e-9, atol=1e-9 * np.abs(c.invK).max())
    pNK, invpNK = onp.get_pNK_test_obs(c.ls, c.sigmas, c.sigma0s, c.nhyp, c.X, c.test_xs)
    np.testing.assert_allclose(pNK, c["pNK"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(invpNK, c.invpNK, rtol=1e-8, atol=1e-10 * np.abs(c.invpNK).max())
    for i in range(c.nhyp):
        m, cv = onp.compute_mean_var_f(c["test_xs0"], c.X, c.Y, c.ls[i], c.sigmas[i], c.sigma0s[i],
                                       c.invK[i], fullcov=True)
        np.testing.assert_allclose(m.reshape(-1), c["test_mean0"][i], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(cv, c["test_cov0"][i], rtol=1e-10, atol=1e-12)
        m, v = onp.compute_mean_var_f(c.x, c.X, c.Y, c.ls[i], c.sigmas[i], c.sigma0s[i], c.invK[i])
        np.testing.assert_allclose(m.reshape(-1), c["meanf_x"][i], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(v.reshape(-1), c["varf_x"][i], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("mode", ["sample", "empirical"])
def test_stage_c_multi_hyp(case, tag, mode):
    """utils.get_testidxs_stats (utils.py:1733-1833) incl. nhyp = 2 (intersection of the maximizer sets)."""
    c = case(tag)
    out = onp.get_testidxs_stats(c.nhyp, c["test_xs0"].copy(), c["test_mean0"].copy(), c["test_cov0"].copy(),
                                 mode=mode, nsample=int(c["nsample"]), n_min_sample=int(c["nmin"]),
                                 z=c["z_joint"], transforms=c["transform"])
    for k, o in enumerate(out):
        ref = c["%s_out%d" % (mode, k)]
        assert np.asarray(o).shape == ref.shape
        if ref.dtype == bool:
            np.testing.assert_array_equal(o, ref)
        else:
            np.testing.assert_allclose(o, ref, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("tag", CASES)
def test_ep_moments(case, tag):
    c = case(tag)
    T = c.test_xs.shape[0]
    for i in range(c.nhyp):
        for j in range(T):
            m, cv = onp.approximate_EP(j, c.mean_k[i].reshape(-1, 1), c.cov_k[i], 1e-9, 100)
            np.testing.assert_allclose(m.squeeze(), c["ep_mean"][i, j], rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(cv, c["ep_cov"][i, j], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("tag", CASES)
def test_tes_sp(case, tag):
    """evaluate_sample_mp.mp (evaluate_sample_mp.py:132-336)."""
    c = case(tag)
    a = onp.evaluate_sample_mp(c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.Yn,
                               c.test_xs, c.max_probs, c.samples, c.masks, c.invpNK, niteration=c.I,
                               z=c["sp_z"])
    np.testing.assert_allclose(a, c["sp_acq"], rtol=RT, atol=AT)


@pytest.mark.parametrize("tag", CASES)
def test_tes_sp_batch(case, tag):
    """evaluate_batch_sample_mp.mp (evaluate_batch_sample_mp.py:160-382)."""
    c = case(tag)
    a = onp.evaluate_batch_sample_mp(c.xq, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.n, c.nhyp, c.Yn,
                                     c.test_xs, c.max_probs, c.samples, c.masks, c.invpNK,
                                     niteration=c.I, z=c["bsp_z"])
    np.testing.assert_allclose(a, c["bsp_acq"], rtol=RT, atol=AT)


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("which", ["emp", "ep"])
def test_tes_ep(case, tag, which):
    """evaluate_mp.mp deterministic + stochastic (evaluate_mp.py:116-295), evaluate_mp_lite.mp
    (evaluate_mp_lite.py:116-225) and the statistics of evaluate_mp.py:56-113."""
    c = case(tag)
    pm, pc = c.moments(which)
    args = (c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.Yn, c.test_xs,
            c.max_probs, pm, pc, c.invK, c.invpNK)
    det = onp.evaluate_mp(*args, stochastic=False)
    # the entropy constant is float32-rounded in the reference (Q12): with it the match is to rounding
    np.testing.assert_allclose(det, c["ep_det_%s_acq" % which], rtol=1e-12, atol=1e-14)
    st = onp.evaluate_mp(*args, stochastic=True, niteration=c.I, z=c["ep_stoch_%s_z" % which])
    np.testing.assert_allclose(st, c["ep_stoch_%s_acq" % which], rtol=RT, atol=AT)
    lite = onp.evaluate_mp_lite(c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.test_xs,
                                c.max_probs, pm, pc, c.invpNK)
    np.testing.assert_allclose(lite, c["ep_lite_%s_acq" % which], rtol=RT, atol=AT)
    qm, qv = onp.get_queried_f_stat_given_data_maxidx(c.x, c.ls[0], c.sigmas[0], c.X, c.Y, c.test_xs,
                                                      c.invpNK[0], pm[0], pc[0])
    np.testing.assert_allclose(qm, c["qmean_%s" % which], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(qv, c["qvar_%s" % which], rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("tag", CASES)
def test_tes_ep_batch_is_zero(case, tag):
    """evaluate_batch_mp.mp (evaluate_batch_mp.py:145-331): -(sum p) log(sum p) ~ 0."""
    c = case(tag)
    a = onp.evaluate_batch_mp(c.xq, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.n, c.nhyp, c.Yn, c.test_xs,
                              c.max_probs, c.emp_mean, c.emp_cov, c.invK, c.invpNK, niteration=c.I,
                              z=c["bep_z"])
    ref = c["bep_acq"]
    assert np.all(np.abs(ref) < 1e-12)
    np.testing.assert_allclose(a, ref, rtol=0, atol=1e-12)
    # a NaN candidate comes back NaN from the reference, every other row stays ~0
    nan_ref = c["bep_acq_nan"]
    assert np.isnan(nan_ref[1]) and np.all(np.abs(np.delete(nan_ref, 1)) < 1e-12)


def test_norm_entropy_constant_is_float32_rounded():
    """Q12: utils.py:653-654 -> tf.cast(0.5 * tf.log(2*np.pi), dtype) is a float32 computation."""
    assert onp.NORM_ENTROPY_CONST == float(np.float32(0.5) * np.log(np.float32(2 * np.pi)))
    assert abs(onp.NORM_ENTROPY_CONST - onp.HALF_LOG_2PI) > 4e-8


# ---- Adam refinement of the trusted maximizers: reference graph code on the lazy-graph shim ---------------------
ADAM_CASES = ["branin", "d6", "h2clip"]


@pytest.mark.parametrize("tag", ADAM_CASES)
def test_adam_refinement_matches_reference_graph(golden, tag):
    """oracle adam_refine_maximizers (+ top-k start points) vs utils_for_continuous.sample_xmaxs_fmaxs run unmodified
    (tests/golden/make_golden_adam.py: tf.train.AdamOptimizer restated per TF 1.14, gradients by autograd through the
    reference's own graph)."""
    g = golden("adam")
    G = lambda n: g["%s_%s" % (tag, n)]  # noqa: E731
    theta, W, b, init, ls, sigmas = G("theta"), G("W"), G("b"), G("init"), G("ls"), G("sigmas")
    k, ntrain = int(G("k")), int(G("ntrain"))
    for i in range(theta.shape[0]):
        idx, _ = onp.rff_topk(init, theta[i], W[i], b[i], sigmas[i], k)
        start = init[idx]                                                   # (S, k, d)
        np.testing.assert_array_equal(start, G("start")[i])
        xs, fv = onp.adam_refine_maximizers(start, theta[i], W[i], b[i], sigmas[i], G("xmin"), G("xmax"), ntrain)
        np.testing.assert_allclose(xs, G("xs")[i], rtol=0, atol=1e-12)
        best = np.argmax(fv, axis=1)
        rows = np.arange(fv.shape[0])
        np.testing.assert_allclose(xs[rows, best], G("max_xs")[i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(fv[rows, best], G("max_fs")[i], rtol=1e-12, atol=1e-13)
        assert np.all(xs >= G("xmin") - 1e-15) and np.all(xs <= G("xmax") + 1e-15)
