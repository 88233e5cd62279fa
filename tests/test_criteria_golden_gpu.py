"""Stage B-D parity against outputs of the REFERENCE's own code: the CUDA path, called through the
reference-named modules (tes_b200.criteria.evaluate_*.mp, tes_b200.utils.*), vs
tests/golden/criteria.npz -- the reference's criteria/*.py + utils.py run unmodified on the numpy
TF/TFP shim with recorded draws (tests/golden/make_golden_criteria.py).

Bar (north_star): acquisition values within 1e-6 relative in fp64 (+1e-9 absolute floor near zero),
selected candidates identical; batch TES_ep (~0 by construction) absolute 1e-12.
"""
import numpy as np
import pytest

from tests.test_oracle_criteria_golden import CASES, Case

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-6, 1e-9


def _close(a, ref):
    a = np.asarray(a)
    np.testing.assert_allclose(a, ref, rtol=RTOL, atol=ATOL)
    assert int(np.argmax(a)) == int(np.argmax(ref))


@pytest.fixture(scope="module")
def case(golden):
    g = golden("criteria")
    return lambda tag: Case(g, tag)


@pytest.mark.parametrize("tag", CASES)
def test_stage_b_against_reference(case, tag):
    """utils.precomputeInvKs, utils.compute_mean_var_f (both branches), get_pNK_test_obs."""
    from tes_b200 import utils
    from tes_b200.criteria import evaluate_mp
    c = case(tag)
    invK = utils.precomputeInvKs(c.d, c.nhyp, c.ls, c.sigmas, c.sigma0s, c.X)
    np.testing.assert_allclose(invK, c.invK, rtol=1e-7, atol=1e-9 * np.abs(c.invK).max())
    pNK, invpNK = evaluate_mp.get_pNK_test_obs(c.ls, c.sigmas, c.sigma0s, c.nhyp, c.X, c.test_xs)
    np.testing.assert_allclose(pNK, c["pNK"], rtol=1e-12, atol=1e-14)
    # the inverse of an ill-conditioned matrix: compare through the residual and to 1e-6 of its scale
    np.testing.assert_allclose(invpNK, c.invpNK, rtol=1e-5, atol=1e-7 * np.abs(c.invpNK).max())
    for i in range(c.nhyp):
        m, cv = utils.compute_mean_var_f(c["test_xs0"], c.X, c.Y, c.ls[i], c.sigmas[i], c.sigma0s[i],
                                         NKInv=c.invK[i], fullcov=True)
        np.testing.assert_allclose(m, c["test_mean0"][i], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(cv, c["test_cov0"][i], rtol=1e-9, atol=1e-11)
        m, v = utils.compute_mean_var_f(c.x, c.X, c.Y, c.ls[i], c.sigmas[i], c.sigma0s[i], NKInv=c.invK[i])
        np.testing.assert_allclose(m, c["meanf_x"][i], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(v, c["varf_x"][i], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("mode", ["sample", "empirical"])
def test_stage_c_against_reference(case, tag, mode):
    """utils.get_testidxs_stats with the reference's draws and colouring factors (nhyp = 1 and 2)."""
    from tes_b200 import utils
    c = case(tag)
    out = utils.get_testidxs_stats(c.nhyp, c["test_xs0"].copy(), c["test_mean0"].copy(), c["test_cov0"].copy(),
                                   mode=mode, nsample=int(c["nsample"]), n_min_sample=int(c["nmin"]),
                                   z=c["z_joint"], transforms=c["transform"])
    for k, o in enumerate(out):
        ref = c["%s_out%d" % (mode, k)]
        o = np.asarray(o)
        assert o.shape == ref.shape
        if ref.dtype == bool:
            np.testing.assert_array_equal(o.astype(bool), ref)
        else:
            np.testing.assert_allclose(o, ref, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("tag", CASES)
def test_ep_moments_against_reference(case, tag):
    """ep.approximate_EP_np outputs for every maximizer (ep.py:73-204)."""
    from tes_b200 import ops
    c = case(tag)
    pm, pc = ops.ep_moments(c.mean_k, c.cov_k, resolution=1e-9, max_niter=100)
    np.testing.assert_allclose(pm.cpu().numpy(), c["ep_mean"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(pc.cpu().numpy(), c["ep_cov"], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("tag", CASES)
def test_tes_sp_against_reference(case, tag):
    from tes_b200.criteria import evaluate_sample_mp
    c = case(tag)
    a = evaluate_sample_mp.mp(c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.Yn, c.test_xs,
                              c.max_probs, c.samples, c.masks, c.invpNK, niteration=c.I, z=c["sp_z"])
    _close(a, c["sp_acq"])
    qm, qv = evaluate_sample_mp.get_queried_f_stat_given_test_samples(
        c.x, c.ls[0], c.sigmas[0], c.sigma0s[0], c.test_xs.shape[0], c.n, c.X, c.Y, c.test_xs,
        c.invpNK[0][:, :c.test_xs.shape[0]], c.invpNK[0][:, c.test_xs.shape[0]:], c.samples[0])
    assert qm.shape == (c.samples.shape[1], c.nx, c.samples.shape[3]) and qv.shape == (c.nx,)


@pytest.mark.parametrize("tag", CASES)
def test_tes_sp_batch_against_reference(case, tag):
    from tes_b200.criteria import evaluate_batch_sample_mp
    c = case(tag)
    a = evaluate_batch_sample_mp.mp(c.xq, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.n, c.nhyp, c.Yn, c.test_xs,
                                    c.max_probs, c.samples, c.masks, c.invpNK, niteration=c.I, z=c["bsp_z"])
    _close(a, c["bsp_acq"])


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("which", ["emp", "ep"])
def test_tes_ep_against_reference(case, tag, which):
    from tes_b200.criteria import evaluate_mp, evaluate_mp_lite
    c = case(tag)
    pm, pc = c.moments(which)
    args = (c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.Yn, c.test_xs, c.max_probs, pm, pc,
            c.invK, c.invpNK)
    _close(evaluate_mp.mp(*args, stochastic=False), c["ep_det_%s_acq" % which])
    _close(evaluate_mp.mp(*args, stochastic=True, niteration=c.I, z=c["ep_stoch_%s_z" % which]),
           c["ep_stoch_%s_acq" % which])
    _close(evaluate_mp_lite.mp(c.x, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.nx, c.n, c.nhyp, c.test_xs,
                               c.max_probs, pm, pc, c.invpNK), c["ep_lite_%s_acq" % which])
    qm, qv = evaluate_mp.get_queried_f_stat_given_data_maxidx(
        c.x, c.ls[0], c.sigmas[0], c.sigma0s[0], c.test_xs.shape[0], c.n, c.X, c.Y, c.test_xs, c.invpNK[0],
        pm[0], pc[0])
    np.testing.assert_allclose(qm, c["qmean_%s" % which], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(qv, c["qvar_%s" % which], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("tag", CASES)
def test_tes_ep_batch_against_reference(case, tag):
    """~0 for every input, NaN for a NaN candidate (evaluate_batch_mp.py:286-331)."""
    from tes_b200.criteria import evaluate_batch_mp
    c = case(tag)
    a = evaluate_batch_mp.mp(c.xq, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.n, c.nhyp, c.Yn, c.test_xs,
                             c.max_probs, c.emp_mean, c.emp_cov, c.invK, c.invpNK, niteration=c.I)
    np.testing.assert_allclose(a, c["bep_acq"], rtol=0, atol=1e-12)
    xq = c.xq.copy()
    xq[1, 0, 0] = np.nan
    a = evaluate_batch_mp.mp(xq, c.ls, c.sigmas, c.sigma0s, c.X, c.Y, c.d, c.n, c.nhyp, c.Yn, c.test_xs,
                             c.max_probs, c.emp_mean, c.emp_cov, c.invK, c.invpNK, niteration=c.I)
    ref = c["bep_acq_nan"]
    assert np.array_equal(np.isnan(a), np.isnan(ref))
    np.testing.assert_allclose(a[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", ["branin", "d6", "h2clip"])
def test_adam_refinement_against_reference(golden, tag):
    """tes_rff_adam_refine_f64 through utils_for_continuous.sample_xmaxs_fmaxs vs the reference's own graph code run on
    the lazy-graph shim (tests/golden/adam.npz)."""
    from tes_b200 import utils_for_continuous as ufc
    g = golden("adam")
    G = lambda n: g["%s_%s" % (tag, n)]  # noqa: E731
    theta, W, b, init, ls, sigmas = G("theta"), G("W"), G("b"), G("init"), G("ls"), G("sigmas")
    nhyp, S, F, d = W.shape
    k, ntrain = int(G("k")), int(G("ntrain"))
    inits, max_xs, max_fs, _, xs = ufc.sample_xmaxs_fmaxs(d, nhyp, S, F, ls, sigmas, np.full(nhyp, 1e-4), theta, W, b, init,
                                                         k, xmin=G("xmin"), xmax=G("xmax"), ntrain=ntrain, get_xs=True)
    np.testing.assert_array_equal(inits, G("start"))
    np.testing.assert_allclose(xs, G("xs"), rtol=0, atol=1e-10)
    np.testing.assert_allclose(max_xs, G("max_xs"), rtol=0, atol=1e-10)
    np.testing.assert_allclose(max_fs, G("max_fs"), rtol=1e-10, atol=1e-11)
