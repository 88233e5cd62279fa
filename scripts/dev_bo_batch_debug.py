"""Device bo_batch loop vs the same loop on the oracle backend: print where the traces part."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pytest
from tests import test_bo_loop_gpu as T
from tests.conftest import GOLDEN
golden = lambda name: np.load(os.path.join(GOLDEN, name + ".npz"))
from tes_b200 import bo_batch
import tes_b200.criteria.evaluate_batch_sample_mp as ebs
q = int(sys.argv[1]) if len(sys.argv) > 1 else 1
f, xs, l, sigma, sigma0 = T._branin(golden)
args = dict(criterion="sftl", mode="sample", nquery=2, nrun=1, ninit=3, nmax=8, nfeature=64, nsto=2, ntrain=25,
            nysample=4, batchsize=q, ntestsample=300, seed=1, transform_fn=T._stable_factor(l, sigma, sigma0))
ta, tb = [], []
out = bo_batch.run(f, xs, 0.0, 1.0, l, sigma, sigma0, trace=ta, **args)
mp = pytest.MonkeyPatch()
ob = T._OracleBackend()
for name in ("precomputeInvKs", "compute_mean_f", "compute_mean_var_f", "get_testidxs_stats"):
    mp.setattr(bo_batch.utils, name, getattr(ob, name))
mp.setattr(bo_batch.optfunc, "draw_random_init_weights_features_np", ob.draw_random_init_weights_features_np)
mp.setattr(bo_batch.utils_for_continuous, "sample_xmaxs_fmaxs", ob.sample_xmaxs_fmaxs)
mp.setattr(ebs, "mp", ob.batch_sample_mp)
mp.setattr(bo_batch, "dev", lambda a: a)
mp.setattr(bo_batch.evaluate_batch_mp_mod(), "get_pNK_test_obs", lambda ls, sg, s0, nh, X, tx: T.onp.get_pNK_test_obs(ls, sg, s0, nh, X, tx))
ref = bo_batch.run(f, xs, 0.0, 1.0, l, sigma, sigma0, trace=tb, **args)
for i, (a, b) in enumerate(zip(ta, tb)):
    print("query", i)
    for k in a:
        if k in b:
            x, y = np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float)
            print("  %-10s shape %s/%s maxdiff %s" % (k, x.shape, y.shape, np.abs(x - y).max() if x.shape == y.shape else "SHAPE"))
    if "start_vals" in a:
        print("   dev start_vals", a["start_vals"], "\n   ref start_vals", b.get("start_vals"))
        print("   dev qx", a["qx"], a["qv"], "\n   ref qx", b.get("qx"), b.get("qv"))
