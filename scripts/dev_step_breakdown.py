import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tes_b200 import acquisition, utils, ops
from tes_b200.device import dev
from tes_b200.criteria import evaluate_mp
torch.cuda.set_device(0)
wl = bench.make_workload("c3", 0, 1)
hyp = wl["hyp"]
devt = {k: torch.from_numpy(wl[k]).cuda() for k in ["cand", "X", "Y", "theta", "W", "b"]}
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = T()
    idx, val, pts = acquisition.trusted_maximizers(devt["cand"], devt["W"], devt["b"], devt["theta"], hyp["sigma"])
    t1 = T()
    test_xs = acquisition.select_test_points(pts[:, 0], 1e-3, 128)
    t2 = T()
    d = 6; ls, sg, s0 = hyp["l"].reshape(1, d), np.array([hyp["sigma"]]), np.array([hyp["sigma0"]])
    invK = utils.precomputeInvKs(d, 1, ls, sg, s0, devt["X"])
    mean_t, cov_t = utils.compute_mean_var_f(dev(test_xs), devt["X"], devt["Y"], ls[0], hyp["sigma"], hyp["sigma0"], invK[0], fullcov=True)
    T0 = test_xs.shape[0]
    m_np, c_np = mean_t.reshape(1, T0).cpu().numpy(), cov_t.reshape(1, T0, T0).cpu().numpy()
    t3 = T()
    z = np.random.RandomState(rep).normal(size=(1, 8192, T0, 1))
    t4 = T()
    out = utils.get_testidxs_stats(1, test_xs, m_np, c_np, mode="empirical", nsample=8192, n_min_sample=2, z=z)
    t5 = T()
    _, invpNK = evaluate_mp.get_pNK_test_obs(ls, sg, s0, 1, devt["X"], dev(out[0]))
    t6 = T()
    acq = evaluate_mp.mp(devt["cand"], ls, sg, s0, devt["X"], devt["Y"], d, 10**6, 64, 1, 10, dev(out[0]), dev(out[1]), dev(out[4]), dev(out[5]), invK, invpNK)
    t7 = T()
    i, v = ops.argmax(acq); i.cpu()
    t8 = T()
    print("A %.1f | select %.1f | B(invK,mean/cov) %.1f | z-gen %.1f | stageC host %.1f | pNK inv %.1f | D %.1f | argmax %.1f | total %.1f ms" % tuple(1e3 * x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6, t8-t7, t8-t0)))
