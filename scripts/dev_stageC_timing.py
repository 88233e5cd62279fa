"""Timing of the stage-C pieces at C3 sizes: python dev_stageC_timing.py [T] [nsample]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops, utils
torch.cuda.set_device(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rs = np.random.RandomState(0)
xs = rs.rand(T, 6); X = rs.rand(64, 6); Y = rs.randn(64, 1)
l = np.array([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949])
invK = utils.precomputeInvKs(6, 1, l.reshape(1, 6), np.array([1.423]), np.array([1e-4]), torch.from_numpy(X).cuda())
mean, cov = utils.compute_mean_var_f(torch.from_numpy(xs).cuda(), torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(), l, 1.423, 1e-4, invK[0], fullcov=True)
def tm(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (name, (time.perf_counter() - t0) / n * 1e3), flush=True)
    return r
e, v, sw = ops.sym_eig(cov, return_sweeps=True); print("sweeps", int(sw[0]))
A = tm("color_factor (jacobi)", lambda: ops.color_factor(cov))
z = tm("normal_fill", lambda: ops.normal_fill(1, 0, 5, ns * T).reshape(ns, T))
samples = tm("gemm z A^T", lambda: ops.gemm(z, A, transb=True))
masked, maxidx, counts = tm("joint_maxidx", lambda: ops.joint_maxidx(samples.clone(), mean.reshape(-1), True))
order = tm("sort", lambda: torch.sort(maxidx, stable=True).indices)
tm("counts D2H", lambda: counts.cpu())
tm("get_testidxs_stats empirical", lambda: utils.get_testidxs_stats(1, torch.from_numpy(xs).cuda(), mean.reshape(1, T), cov.reshape(1, T, T), mode="empirical", nsample=ns, seed=3))
tm("get_testidxs_stats sample", lambda: utils.get_testidxs_stats(1, torch.from_numpy(xs).cuda(), mean.reshape(1, T), cov.reshape(1, T, T), mode="sample", nsample=ns, seed=3))
pts = torch.from_numpy(xs[rs.randint(0, T, 1024)]).cuda()
from tes_b200 import acquisition
tm("select_test_points S=1024", lambda: acquisition.select_test_points(pts, 1e-3, 128))
