"""Device EP timing vs the host restatement: python dev_ep_device_timing.py T [T ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops
from oracle import tes_oracle_np as onp
torch.cuda.set_device(0)
for T in [int(a) for a in sys.argv[1:]] or [8, 16, 49, 125]:
    rs = np.random.RandomState(T)
    xs = rs.rand(T, 2) * 4.0; X = rs.rand(5, 2) * 4.0; Y = rs.randn(5, 1); l = np.array([3.0, 5.0])
    invK = onp.precomputeInvKs(l.reshape(1, 2), np.array([1.0]), np.array([1e-2]), X)
    m, c = onp.compute_mean_var_f(xs, X, Y, l, 1.0, 1e-2, invK[0], fullcov=True)
    c = c + 1e-6 * np.eye(T)
    md, cd = torch.from_numpy(m.reshape(1, T)).cuda(), torch.from_numpy(c.reshape(1, T, T)).cuda()
    ops.ep_moments(md, cd, 1e-9, 100); torch.cuda.synchronize()
    t0 = time.perf_counter(); pm, pc, nit = ops.ep_moments(md, cd, 1e-9, 100, return_niter=True); torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    nh = min(T, 4)
    t0 = time.perf_counter()
    for j in range(nh): onp.approximate_EP(j, m.reshape(-1, 1), c, 1e-9, 100)
    t_host = (time.perf_counter() - t0) / nh * T
    print("T=%d: device (all %d maximizers) %.2f ms, iterations %d..%d; host numpy (extrapolated from %d) %.1f ms -> %.0fx" % (
        T, T, t_dev * 1e3, int(nit.min()), int(nit.max()), nh, t_host * 1e3, t_host / t_dev), flush=True)
