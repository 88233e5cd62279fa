"""Analytic cross-checks of the numpy oracle for the parts of the reference that no reference test
pins (the five TF criteria, EP): parity for those is otherwise UNPINNED."""
import numpy as np
import pytest
import scipy.integrate as spi
import scipy.stats as spst

from oracle import tes_oracle_np as onp
from tests._problems import BRANIN, PH, make_problem


def _moments(pr):
    p = pr["max_probs"][0]
    keep = p > 0
    qm, qv = onp.get_queried_f_stat_given_data_maxidx(pr["x"], pr["l"], pr["sigma"], pr["X"], pr["Y"], pr["test_xs"],
                                                      pr["invpNK"][0], pr["v0"][0][keep], pr["v1"][0][keep])
    return p[keep], qm, qv


def test_lite_is_pairwise_kl():
    pr = make_problem(seed=1, nx=11, mode="empirical", nsample=400)
    p, qm, qv = _moments(pr)
    v = qv + pr["sigma0"]
    ref = np.zeros(pr["nx"])
    for a in range(len(p)):
        for b in range(len(p)):
            kl = 0.5 * np.log(v[a] / v[b]) + (v[b] + (qm[b] - qm[a]) ** 2) / (2 * v[a]) - 0.5  # KL(N_b || N_a)
            ref += p[a] * p[b] * kl
    got = onp.evaluate_mp_lite(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"], pr["n"],
                               1, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)
    assert np.all(got > -1e-12)


def test_deterministic_matches_hand_formula_with_reduce_mean_quirk():
    pr = make_problem(seed=2, nx=9, mode="empirical", nsample=400, hyp=PH, n=15)
    p, qm, qv = _moments(pr)
    _, varf = onp.compute_mean_var_f(pr["x"], pr["X"], pr["Y"], pr["l"], pr["sigma"], pr["sigma0"], pr["invK"][0])
    # analytic normal entropy, shifted by the reference's float32-rounded constant (Q12, utils.py:653-654)
    H = lambda s: 0.5 * np.log(2 * np.pi * np.e * s ** 2) + (onp.NORM_ENTROPY_CONST - onp.HALF_LOG_2PI)
    ref = H(np.sqrt(varf + pr["sigma0"])) - np.sum(p[:, None] * H(np.sqrt(qv + pr["sigma0"])), axis=0) / len(p)  # Q4
    got = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"], pr["n"], 1, 10,
                          pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)


def test_stochastic_estimator_converges_to_mixture_entropy_gap():
    pr = make_problem(seed=3, nx=2, mode="empirical", nsample=400)
    p, qm, qv = _moments(pr)
    sd = np.sqrt(qv + pr["sigma0"])
    I, Yn = 40, 500
    z = np.random.RandomState(0).normal(size=(1, I, Yn, len(p), pr["nx"]))
    got = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"], pr["n"], 1,
                          Yn, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"],
                          stochastic=True, niteration=I, z=z)
    for x in range(pr["nx"]):
        mix = lambda y: np.sum(p * spst.norm.pdf(y, qm[:, x], sd[:, x]))
        lo, hi = (qm[:, x] - 10 * sd[:, x]).min(), (qm[:, x] + 10 * sd[:, x]).max()
        Hmix = -spi.quad(lambda y: mix(y) * np.log(max(mix(y), 1e-300)), lo, hi, limit=400)[0]
        Hcond = np.sum(p * 0.5 * np.log(2 * np.pi * np.e * sd[:, x] ** 2))
        assert abs(got[x] - (Hmix - Hcond)) < 0.02  # MC error ~ 1/sqrt(20000 * S')


def test_batch_tes_ep_is_identically_zero():
    pr = make_problem(seed=4, nx=5, mode="empirical", nsample=400)
    p = pr["max_probs"][0]
    Sp = int((p > 0).sum())
    rs = np.random.RandomState(1)
    xb = rs.rand(5, 3, 2)
    z = rs.normal(size=(1, 2, 4, Sp, 5, 3))
    got = onp.evaluate_batch_mp(xb, pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["n"], 1, 4,
                                pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invK"], pr["invpNK"],
                                niteration=2, z=z)
    s = p[p > 0].sum()
    assert np.max(np.abs(got - (-s * np.log(s)))) < 1e-12 and np.max(np.abs(got)) < 1e-12


def test_tes_sp_with_one_sample_per_maximizer_equals_tes_ep_with_zero_cov():
    """K = 1, all masks valid: every mixture is a single Gaussian N(M m_j + b, Kq + sigma0), which is TES_ep with
    post_mean_j = m_j and post_cov_j = 0 -- ties the two criterion families together on identical draws."""
    pr = make_problem(seed=5, nx=6, mode="empirical", nsample=400)
    p = pr["max_probs"]
    T = pr["T"]
    rs = np.random.RandomState(2)
    m = rs.randn(1, T, T) * 0.3
    I, Yn = 3, 5
    z = rs.normal(size=(1, I, Yn, T, pr["nx"]))
    ep_val = onp.evaluate_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"], pr["n"], 1,
                             Yn, pr["test_xs"], p, m, np.zeros((1, T, T, T)), pr["invK"], pr["invpNK"], stochastic=True,
                             niteration=I, z=z)
    sp_val = onp.evaluate_sample_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"],
                                    pr["n"], 1, Yn, pr["test_xs"], p, m[..., None], np.ones((1, T, 1), dtype=bool),
                                    pr["invpNK"], niteration=I, z=z[..., None])
    np.testing.assert_allclose(sp_val, ep_val, rtol=1e-9, atol=1e-12)


def test_batch_tes_sp_with_q1_equals_tes_sp():
    pr = make_problem(seed=6, nx=4, mode="sample", nsample=120)
    p = pr["max_probs"][0]
    Sp, K = int((p > 0).sum()), pr["v0"].shape[-1]
    I, Yn = 2, 3
    z = np.random.RandomState(3).normal(size=(1, I, Yn, Sp, pr["nx"], K))
    a = onp.evaluate_sample_mp(pr["x"], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2, pr["nx"], pr["n"], 1,
                               Yn, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"], niteration=I, z=z)
    b = onp.evaluate_batch_sample_mp(pr["x"][:, None, :], pr["ls"], pr["sigmas"], pr["sigma0s"], pr["X"], pr["Y"], 2,
                                     pr["n"], 1, Yn, pr["test_xs"], pr["max_probs"], pr["v0"], pr["v1"], pr["invpNK"],
                                     niteration=I, z=z[..., None])
    np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-11)


def test_ep_two_points_is_exact_truncated_moments():
    """T = 2: one site, the tilted distribution is N(mu,Sigma) truncated to c^T f >= 0, whose first two moments
    have a closed form -- EP's fixed point must reproduce them."""
    mu = np.array([[0.3], [0.7]])
    Sig = np.array([[0.5, 0.2], [0.2, 0.8]])
    for j in (0, 1):
        c = np.array([[1.0], [-1.0]]) * (1 if j == 0 else -1)
        s = np.sqrt((c.T @ Sig @ c).item())
        beta = (c.T @ mu).item() / s
        lam = spst.norm.pdf(beta) / spst.norm.cdf(beta)
        mean = mu + Sig @ c * lam / s
        cov = Sig - (Sig @ c @ c.T @ Sig) * lam * (lam + beta) / s ** 2
        m, C = onp.approximate_EP(j, mu, Sig, 1e-14, 200)
        np.testing.assert_allclose(m, mean, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(C, cov, rtol=1e-7, atol=1e-9)


def test_chol2inv_and_pnk_identities():
    pr = make_problem(seed=7, n=9, T=5)
    NK = onp.computeNKmm(pr["X"], pr["l"], pr["sigma"], pr["sigma0"])
    np.testing.assert_allclose(onp.chol2inv(NK) @ NK, np.eye(9), atol=1e-8)
    pNK, inv = onp.get_pNK_test_obs(pr["ls"], pr["sigmas"], pr["sigma0s"], 1, pr["X"], pr["test_xs"])
    Tk = pr["T"]
    assert np.allclose(np.diag(pNK[0])[:Tk], pr["sigma"]) and np.allclose(np.diag(pNK[0])[Tk:], pr["sigma"] + pr["sigma0"])
    np.testing.assert_allclose(inv[0] @ pNK[0], np.eye(Tk + 9), atol=1e-6)
