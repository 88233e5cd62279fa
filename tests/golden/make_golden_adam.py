"""Golden fixtures for the Adam refinement of the trusted maximizers, produced by running the REFERENCE's own
utils_for_continuous.sample_xmaxs_fmaxs (utils_for_continuous.py:138-252) UNMODIFIED on the lazy-graph TF-1.14 shim
(oracle/tf_shim/graph.py: variables, assign, tf.train.AdamOptimizer with torch-autograd gradients through the
reference's graph, Session.run), the way the drivers use it (bo_batch.py:460-476, :676-690): build the graph,
sess.run(assigns), `ntrain` x sess.run(trains), sess.run([max_xs, max_fs, xs_all]).  Build container only.

    python tests/golden/make_golden_adam.py        ->  tests/golden/adam.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # tag: (seed, d, nhyp, S, F, n_init, k, ntrain, xmin, xmax, l, sigma)
    "branin": (1, 2, 1, 6, 64, 120, 3, 40, 0.0, 1.0, np.array([25.5866, 6.3735]), 0.589),
    "d6": (2, 6, 1, 4, 96, 200, 2, 25, 0.0, 1.0, np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949]), 1.423),
    "h2clip": (3, 3, 2, 3, 48, 60, 3, 60, np.array([0.0, 0.2, 0.0]), np.array([1.0, 0.8, 0.5]),
               np.array([12.0, 3.0, 7.5]), 0.8),       # two hyper-parameter samples, a box that clips
}


def main():
    rl.use_graph_shim()
    import torch
    ufc = rl.load("utils_for_continuous")
    tf = sys.modules["tensorflow"]
    out = {}
    for tag, (seed, d, nhyp, S, F, n_init, k, ntrain, xmin, xmax, l, sigma) in CASES.items():
        rs = np.random.RandomState(500 + seed)
        ls = np.stack([l * (1.0 + 0.1 * i) for i in range(nhyp)])
        sigmas = np.array([sigma * (1.0 + 0.05 * i) for i in range(nhyp)])
        sigma0s = np.full(nhyp, 1e-4)
        W = rs.randn(nhyp, S, F, d) * np.sqrt(ls)[:, None, None, :]
        b = 2.0 * np.pi * rs.rand(nhyp, S, F, 1)
        theta = rs.randn(nhyp, S, F, 1)
        lo = np.broadcast_to(np.asarray(xmin, dtype=np.float64), (d,))
        hi = np.broadcast_to(np.asarray(xmax, dtype=np.float64), (d,))
        init = lo + rs.rand(n_init, d) * (hi - lo)
        tf.reset_default_graph()
        with rl.quiet():
            assigns, trains, max_xs, max_fs, xs_all = ufc.sample_xmaxs_fmaxs(
                d, nhyp, S, F, torch.from_numpy(ls), torch.from_numpy(sigmas), torch.from_numpy(sigma0s),
                torch.from_numpy(theta), torch.from_numpy(W), torch.from_numpy(b), torch.from_numpy(init), k, xmin, xmax,
                dtype=tf.float64, get_xs=True, name="golden_" + tag)
            sess = tf.Session()
            sess.run(assigns)
            start = sess.run(xs_all)                                  # the top-k start points (before any step)
            for _ in range(ntrain):
                sess.run(trains)
            mx, mf, xs = sess.run([max_xs, max_fs, xs_all])
        P = lambda n: "%s_%s" % (tag, n)  # noqa: E731
        out.update({P("ls"): ls, P("sigmas"): sigmas, P("theta"): theta, P("W"): W, P("b"): b, P("init"): init,
                    P("k"): np.int64(k), P("ntrain"): np.int64(ntrain), P("xmin"): lo.copy(), P("xmax"): hi.copy(),
                    P("start"): start, P("max_xs"): mx, P("max_fs"): mf, P("xs"): xs})
        print(tag, "start", start.shape, "max_fs", mf.reshape(-1)[:4], "moved", float(np.abs(xs - start).max()))
    path = os.path.join(OUT, "adam.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    if not rl.reference_available():
        sys.exit("reference tree not available; fixtures are committed, nothing to do")
    main()
