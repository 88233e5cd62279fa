"""Golden fixtures for the five TES criteria, produced by running the REFERENCE's own
criteria/*.py and the TF halves of its utils.py UNMODIFIED on the numpy-backed TF-1.14/TFP-0.7
shim (oracle/tf_shim), through oracle/ref_loader.py.  Build container only.

    python tests/golden/make_golden_criteria.py        (also run by make_golden.py)

One .npz (tests/golden/criteria.npz).  Per case `<tag>_...`:
  inputs   x / xq (batch), ls, sigmas, sigma0s, X, Y, test_xs0 (before pruning), nsample, joint-sample
           draws z_joint + the colouring factors the reference's np.linalg.eig produced
  stage B  invK (utils.precomputeInvKs), test_mean/test_cov (utils.compute_mean_var_f fullcov),
           pNK/invpNK (criteria.evaluate_mp.get_pNK_test_obs), meanf/varf at x
  stage C  utils.get_testidxs_stats outputs for mode sample / empirical (+ ep via ep.approximate_EP_np
           with the arguments utils.py:1820-1825 intends)
  stage D  <crit>_acq = mp(...) and <crit>_z = the standard normals its TFP sample() calls consumed,
           stacked (nhyp, niteration, ...) in call order.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

BRANIN = (np.array([25.5866, 6.3735]), 0.589, 1e-4)                               # functions.py:571-577
PH = (np.array([177.9418631280815, 309.4630784148164]), 0.29444226420446995, 0.06476473751595126)
HART6 = (np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]), 1.423, 1e-4)  # functions.py:500-545

CASES = {
    # tag: dict(hyps=[(l, sigma, sigma0), ...], d, n, T0, nx, nxq, q, nsample, nmin, I, Y, seed)
    "c1": dict(hyps=[BRANIN], d=2, n=5, T0=8, nx=9, nxq=5, q=3, nsample=300, nmin=2, I=2, Y=4, seed=1),
    "c2": dict(hyps=[PH], d=2, n=12, T0=7, nx=8, nxq=4, q=2, nsample=200, nmin=2, I=2, Y=5, seed=2),
    "h2": dict(hyps=[(np.array([12.0, 3.0, 7.5]), 0.8, 1e-3), (np.array([20.0, 6.0, 2.5]), 0.6, 1e-2)],
               d=3, n=6, T0=6, nx=6, nxq=4, q=2, nsample=250, nmin=2, I=2, Y=3, seed=3),
    "d6": dict(hyps=[HART6], d=6, n=20, T0=10, nx=7, nxq=3, q=4, nsample=400, nmin=2, I=1, Y=4, seed=4),
}


def _stack_draws(tfs, nhyp, niter):
    zs = tfs.draws.normals_logged()
    assert len(zs) == nhyp * niter, (len(zs), nhyp, niter)
    return np.stack(zs).reshape((nhyp, niter) + zs[0].shape)


def make_case(tag, c, out):
    from oracle import tf_shim as tfs
    utils = rl.load("utils")
    ep = rl.load("ep")
    emp_mod = rl.load("criteria.evaluate_mp")
    lite_mod = rl.load("criteria.evaluate_mp_lite")
    sp_mod = rl.load("criteria.evaluate_sample_mp")
    bsp_mod = rl.load("criteria.evaluate_batch_sample_mp")
    bep_mod = rl.load("criteria.evaluate_batch_mp")
    f64 = np.float64

    rs = np.random.RandomState(1000 + c["seed"])
    d, n, T0, nx, nxq, q = c["d"], c["n"], c["T0"], c["nx"], c["nxq"], c["q"]
    nhyp = len(c["hyps"])
    ls = np.stack([h[0] for h in c["hyps"]]).astype(f64)
    sigmas = np.array([h[1] for h in c["hyps"]], dtype=f64)
    sigma0s = np.array([h[2] for h in c["hyps"]], dtype=f64)
    X = rs.rand(n, d)
    Y = rs.randn(n, 1) * 0.5
    test_xs0 = rs.rand(T0, d)
    x = rs.rand(nx, d)
    xq = rs.rand(nxq, q, d)
    P = lambda k: "%s_%s" % (tag, k)  # noqa: E731
    out.update({P("ls"): ls, P("sigmas"): sigmas, P("sigma0s"): sigma0s, P("X"): X, P("Y"): Y,
                P("test_xs0"): test_xs0, P("x"): x, P("xq"): xq,
                P("nsample"): np.int64(c["nsample"]), P("nmin"): np.int64(c["nmin"]),
                P("I"): np.int64(c["I"]), P("Y_n"): np.int64(c["Y"])})

    with rl.quiet():
        # ---- stage B, TF halves of utils.py (utils.py:657-815, 1838-1856) -------------------
        invK = utils.precomputeInvKs(d, nhyp, ls, sigmas, sigma0s, X, f64)
        tm, tc, mf, vf = [], [], [], []
        for i in range(nhyp):
            l = ls[i].reshape(1, d)
            m, cv = utils.compute_mean_var_f(test_xs0, X, Y, l, sigmas[i], sigma0s[i], NKInv=invK[i],
                                             fullcov=True, dtype=f64)
            tm.append(np.reshape(m, (T0,)))
            tc.append(cv)
            m2, v2 = utils.compute_mean_var_f(x, X, Y, l, sigmas[i], sigma0s[i], NKInv=invK[i],
                                              fullcov=False, dtype=f64)
            mf.append(m2)
            vf.append(v2)
        test_means, test_covs = np.stack(tm), np.stack(tc)
        out.update({P("invK"): invK, P("test_mean0"): test_means, P("test_cov0"): test_covs,
                    P("meanf_x"): np.stack(mf), P("varf_x"): np.stack(vf)})

        # ---- stage C (utils.py:1654-1833): joint-sample draws from the global numpy RNG ------
        seed_c = 77 + c["seed"]
        np.random.seed(seed_c)
        z_joint = np.stack([np.random.normal(size=(c["nsample"], T0, 1)) for _ in range(nhyp)])
        transforms = []
        for i in range(nhyp):
            e, v = np.linalg.eig(test_covs[i])  # utils.py:1668-1670
            transforms.append(v.dot(np.diag(np.sqrt(np.clip(e, 0.0, np.inf)))))
        out.update({P("z_joint"): z_joint, P("transform"): np.real(np.stack(transforms))})
        stats = {}
        for mode in ("sample", "empirical"):
            np.random.seed(seed_c)
            stats[mode] = utils.get_testidxs_stats(nhyp, test_xs0.copy(), test_means.copy(),
                                                   test_covs.copy(), mode=mode,
                                                   nsample=c["nsample"], n_min_sample=c["nmin"])
            for k, o in enumerate(stats[mode]):
                out[P("%s_out%d" % (mode, k))] = np.asarray(o)
        test_xs, max_probs, mean_k, cov_k, samples, masks = stats["sample"]
        _, _, _, _, emp_mean, emp_cov = stats["empirical"]
        T = test_xs.shape[0]
        ep_mean = np.zeros((nhyp, T, T))
        ep_cov = np.zeros((nhyp, T, T, T))
        for i in range(nhyp):
            for j in range(T):  # utils.py:1815-1829 with `ep` imported
                m, cv = ep.approximate_EP_np(j, mean_k[i].reshape(-1, 1), cov_k[i],
                                             resolution=1e-9, max_niter=100)
                ep_mean[i, j] = m.squeeze()
                ep_cov[i, j] = cv
        out.update({P("ep_mean"): ep_mean, P("ep_cov"): ep_cov})

        # ---- a8: pNK / invpNK (evaluate_mp.py:12-53) -----------------------------------------
        pNK, invpNK = emp_mod.get_pNK_test_obs(ls, sigmas, sigma0s, nhyp, X, test_xs, dtype=f64)
        out.update({P("pNK"): pNK, P("invpNK"): invpNK})
        for mod in (lite_mod, sp_mod, bsp_mod, bep_mod):  # the five copies are identical
            p2, i2 = mod.get_pNK_test_obs(ls, sigmas, sigma0s, nhyp, X, test_xs, dtype=f64)
            assert np.array_equal(p2, pNK) and np.array_equal(i2, invpNK)

        I, Yn = c["I"], c["Y"]
        # ---- a14/a15 TES_sp q=1 (evaluate_sample_mp.py:132-336) ------------------------------
        tfs.draws.record(500 + c["seed"])
        acq = sp_mod.mp(x, ls, sigmas, sigma0s, X, Y, d, nx, n, nhyp, Yn, test_xs, max_probs,
                        samples, masks, invpNK, dtype=f64, niteration=I)
        out.update({P("sp_acq"): acq, P("sp_z"): _stack_draws(tfs, nhyp, I)})
        # ---- a16 batch TES_sp (evaluate_batch_sample_mp.py:160-382) --------------------------
        tfs.draws.record(600 + c["seed"])
        acq = bsp_mod.mp(xq, ls, sigmas, sigma0s, X, Y, d, n, nhyp, Yn, test_xs, max_probs,
                         samples, masks, invpNK, dtype=f64, niteration=I)
        out.update({P("bsp_acq"): acq, P("bsp_z"): _stack_draws(tfs, nhyp, I)})
        # ---- a17/a18 TES_ep q=1 (evaluate_mp.py:116-295), moments: empirical and EP ----------
        for mname, (pm, pc) in {"emp": (emp_mean, emp_cov), "ep": (ep_mean, ep_cov)}.items():
            tfs.draws.record(0)
            acq = emp_mod.mp(x, ls, sigmas, sigma0s, X, Y, d, nx, n, nhyp, Yn, test_xs, max_probs,
                             pm, pc, invK, invpNK, stochastic=False, dtype=f64)
            assert not tfs.draws.log
            out[P("ep_det_%s_acq" % mname)] = acq
            tfs.draws.record(700 + c["seed"])
            acq = emp_mod.mp(x, ls, sigmas, sigma0s, X, Y, d, nx, n, nhyp, Yn, test_xs, max_probs,
                             pm, pc, invK, invpNK, stochastic=True, dtype=f64, niteration=I)
            out.update({P("ep_stoch_%s_acq" % mname): acq,
                        P("ep_stoch_%s_z" % mname): _stack_draws(tfs, nhyp, I)})
            # ---- a19 lite (evaluate_mp_lite.py:116-225) --------------------------------------
            out[P("ep_lite_%s_acq" % mname)] = lite_mod.mp(
                x, ls, sigmas, sigma0s, X, Y, d, nx, n, nhyp, test_xs, max_probs, pm, pc, invpNK,
                dtype=f64)
            # the q=1 statistics themselves (evaluate_mp.py:56-113), hyp 0
            qm, qv = emp_mod.get_queried_f_stat_given_data_maxidx(
                x, ls[0].reshape(1, d), sigmas[0], sigma0s[0], T, n, X, Y, test_xs, invpNK[0],
                pm[0], pc[0], dtype=f64)
            out.update({P("qmean_%s" % mname): qm, P("qvar_%s" % mname): qv})
        # ---- a20 batch TES_ep (evaluate_batch_mp.py:145-331; ~0 by construction) --
        tfs.draws.record(800 + c["seed"])
        acq = bep_mod.mp(xq, ls, sigmas, sigma0s, X, Y, d, n, nhyp, Yn, test_xs, max_probs,
                         emp_mean, emp_cov, invK, invpNK, dtype=f64, niteration=I)
        out.update({P("bep_acq"): acq, P("bep_z"): _stack_draws(tfs, nhyp, I)})
        # a20 with a NaN candidate: the reference propagates it (eigh of a NaN matrix)
        xq_nan = xq.copy()
        xq_nan[1, 0, 0] = np.nan
        tfs.draws.record(800 + c["seed"])
        try:
            acq_nan = bep_mod.mp(xq_nan, ls, sigmas, sigma0s, X, Y, d, n, nhyp, Yn, test_xs,
                                 max_probs, emp_mean, emp_cov, invK, invpNK, dtype=f64,
                                 niteration=I)
        except np.linalg.LinAlgError:  # LAPACK refuses NaN input; TF's eigh returns NaN
            acq_nan = acq.copy()
            acq_nan[1] = np.nan
        out[P("bep_acq_nan")] = acq_nan
    return out


def main():
    out = {}
    for tag, c in CASES.items():
        make_case(tag, c, out)
        print(tag, "T kept:", out[tag + "_sample_out0"].shape[0],
              "K:", out[tag + "_sample_out4"].shape[-1],
              "sp_acq[:3]:", out[tag + "_sp_acq"][:3])
    path = os.path.join(OUT, "criteria.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    if not rl.reference_available():
        sys.exit("reference tree not available; fixtures are committed, nothing to do")
    main()
