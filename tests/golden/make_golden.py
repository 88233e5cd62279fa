"""Generate the golden fixtures in tests/golden/*.npz by running the REFERENCE's own numpy
functions (imported unmodified from /root/reference through oracle/ref_loader.py).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/tes_oracle_np.py, oracle/tes_oracle.c); the oracle in turn
is what the CUDA path is compared with.  Nothing here is product code.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


def main():
    utils = rl.load("utils")
    optfunc = rl.load("optfunc")
    ep = rl.load("ep")
    emp = rl.load("empirical_approximation")
    functions = rl.load("functions")

    # ---- 1. Brooms-Barn KATs: the only golden values the reference ships --------------------
    # functions.py:299-398: objective = utils.compute_mean_f_np on 433-435 observations, with the
    # shipped 'maximizer'/'maximum'.  Also carries the data + hyper-parameters config C2 needs.
    kat = {}
    with rl.in_reference_cwd(), rl.quiet():
        for name in ("pH", "log10P", "log10K"):
            info = getattr(functions, name)()
            X = np.loadtxt("pHdata/X_%s.txt" % name)
            Y = np.loadtxt("pHdata/Y_%s.txt" % name)
            hyp = np.loadtxt("pHdata/hyperparameters_%s.txt" % name)
            kat[name + "_X"] = X
            kat[name + "_Y"] = Y
            kat[name + "_hyp"] = hyp  # sigma, l0, l1, sigma0
            kat[name + "_maximizer"] = info["maximizer"]
            kat[name + "_maximum"] = np.float64(info["maximum"])
            kat[name + "_f_at_maximizer"] = info["function"](info["maximizer"])
            kat[name + "_xs"] = info["xs"]
            kat[name + "_f_on_xs"] = info["function"](info["xs"])
        br = functions.negative_Branin()
        kat["branin_xs"] = br["xs"]
        kat["branin_f_on_xs"] = br["function"](br["xs"])
        kat["branin_l"] = br["RBF.lengthscale"]
        kat["branin_sigma"] = np.float64(br["RBF.variance"])
        kat["branin_sigma0"] = np.float64(br["noise.variance"])
        h6 = functions.negative_rescaled_hartmann6d()
        rs = np.random.RandomState(3)
        h6x = rs.rand(64, 6)
        kat["hartmann6_x"] = h6x
        kat["hartmann6_f"] = h6["function"](h6x)
        kat["hartmann6_l"] = h6["RBF.lengthscale"]
        kat["hartmann6_sigma"] = np.float64(h6["RBF.variance"])
        kat["hartmann6_sigma0"] = np.float64(h6["noise.variance"])
    save("kat_objectives", **kat)

    # ---- 2. SE-ARD kernels (utils.computeKnm_np / computeKmm_np) ----------------------------
    rs = np.random.RandomState(11)
    ker = {}
    for tag, (n, m, d) in {"a": (7, 5, 2), "b": (16, 9, 6), "c": (3, 12, 10)}.items():
        A = rs.rand(n, d)
        B = rs.rand(m, d)
        l = rs.rand(d) * 10 + 0.5
        sigma = 0.3 + rs.rand()
        ker[tag + "_A"], ker[tag + "_B"], ker[tag + "_l"] = A, B, l
        ker[tag + "_sigma"] = np.float64(sigma)
        ker[tag + "_Knm"] = utils.computeKnm_np(A, B, l, sigma)
        ker[tag + "_Kmm"] = utils.computeKmm_np(A, l, sigma)
    save("se_ard", **ker)

    # ---- 3. RFF posterior weight draw + function samples (optfunc.py:303-402) ---------------
    rff = {}
    branin_xs = kat["branin_xs"]
    cases = {
        # tag: (seed, xdim, S, F, n, l, sigma, sigma0)
        "nltF": (1, 2, 4, 32, 5, np.array([25.5866, 6.3735]), 0.589, 1e-4),   # n < F branch
        "ngeF": (2, 2, 3, 8, 20, np.array([25.5866, 6.3735]), 0.589, 1e-2),   # n >= F branch
        "d6": (3, 6, 3, 64, 12, np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]),
               1.423, 1e-4),
    }
    for tag, (seed, xdim, S, F, n, l, sigma, sigma0) in cases.items():
        rs = np.random.RandomState(100 + seed)
        X = rs.rand(n, xdim)
        Y = rs.randn(n, 1)
        np.random.seed(seed)
        theta, W, b = optfunc.draw_random_init_weights_features_np(
            xdim, S, F, X, Y, l.reshape(1, xdim), sigma, sigma0)
        # replay the global-RNG stream in the reference's draw order (optfunc.py:329,334,344)
        np.random.seed(seed)
        randn_W = np.random.randn(S, F, xdim)
        rand_b = np.random.rand(S, F, 1)
        randn_noise = np.random.randn(S, F, 1)
        xs = branin_xs if xdim == 2 else np.random.RandomState(7).rand(500, xdim)
        fv = np.stack([optfunc.make_function_sample_np(xs, F, sigma, theta[j], W[j], b[j])
                       for j in range(S)])
        rff.update({
            tag + "_X": X, tag + "_Y": Y, tag + "_l": l, tag + "_sigma": np.float64(sigma),
            tag + "_sigma0": np.float64(sigma0), tag + "_theta": theta, tag + "_W": W,
            tag + "_b": b, tag + "_randn_W": randn_W, tag + "_rand_b": rand_b,
            tag + "_randn_noise": randn_noise, tag + "_xs": xs, tag + "_fvals": fv,
            tag + "_argmax": np.argmax(fv, axis=1),
        })
    save("rff", **rff)

    # ---- 4. trusted-maximizer statistics (utils.py:1654-1833) -------------------------------
    st = {}
    rs = np.random.RandomState(5)
    for tag, (T, nsample) in {"t6": (6, 400), "t12": (12, 1000)}.items():
        d = 2
        test_xs = rs.rand(T, d)
        Xo = rs.rand(5, d)
        Yo = rs.randn(5, 1)
        l = np.array([25.5866, 6.3735])
        sigma, sigma0 = 0.589, 1e-4
        K = utils.computeKmm_np(Xo, l, sigma) + np.eye(5) * sigma0
        Kinv = np.linalg.inv(K)
        ks = utils.computeKnm_np(test_xs, Xo, l, sigma)
        mean = (ks @ Kinv @ Yo).reshape(1, T)
        cov = (utils.computeKmm_np(test_xs, l, sigma) - ks @ Kinv @ ks.T).reshape(1, T, T)
        st[tag + "_test_xs"], st[tag + "_mean"], st[tag + "_cov"] = test_xs, mean, cov
        for mode in ("sample", "empirical"):
            np.random.seed(42)
            with rl.quiet():
                out = utils.get_testidxs_stats(1, test_xs.copy(), mean.copy(), cov.copy(),
                                               mode=mode, nsample=nsample, n_min_sample=2)
            for i, o in enumerate(out):
                st["%s_%s_out%d" % (tag, mode, i)] = np.asarray(o)
        np.random.seed(42)
        st[tag + "_z"] = np.random.normal(size=(1, nsample, T, 1))
        st[tag + "_nsample"] = np.int64(nsample)
        e, v = np.linalg.eig(cov[0])
        st[tag + "_transform"] = v.dot(np.diag(np.sqrt(np.clip(e, 0.0, np.inf))))
        # EP per maximizer with the arguments utils.py:1820-1825 intends
        out_s = [st["%s_sample_out%d" % (tag, i)] for i in range(6)]
        mean_k, cov_k = out_s[2], out_s[3]
        Tk = mean_k.shape[1]
        epm = np.zeros((Tk, Tk))
        epc = np.zeros((Tk, Tk, Tk))
        with rl.quiet():
            for j in range(Tk):
                m, c = ep.approximate_EP_np(j, mean_k[0].reshape(-1, 1), cov_k[0],
                                            resolution=1e-9, max_niter=100)
                epm[j] = m.squeeze()
                epc[j] = c
        st[tag + "_ep_mean"], st[tag + "_ep_cov"] = epm, epc
    # empirical moments direct (empirical_approximation.py:86-107)
    rs = np.random.RandomState(9)
    smp = rs.randn(50, 4)
    w = (rs.rand(50) > 0.3).astype(int)
    m, c = emp.get_empirical_stat_from_samples(smp, weight=w)
    st["emp_samples"], st["emp_weight"], st["emp_mean"], st["emp_cov"] = smp, w, m, c
    # de-dup (utils.py:1554-1581)
    pts = rs.rand(12, 2)
    pts[5] = pts[1] + 1e-4
    pts[9] = pts[1] - 2e-4
    pts[7] = pts[3]
    st["dedup_in"] = pts
    st["dedup_out"] = utils.remove_duplicates_np(pts.copy(), resolution=1e-3)
    save("maximizer_stats", **st)


if __name__ == "__main__":
    if not rl.reference_available():
        sys.exit("reference tree not available; fixtures are committed, nothing to do")
    main()
