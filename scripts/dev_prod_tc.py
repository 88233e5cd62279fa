"""tcgen05 product GEMM (mode 3) vs the mma.sync fp16 mode (2): identical results + timing.  python dev_prod_tc.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops
rs = np.random.RandomState(0)
def mesh(axes):
    xds = np.meshgrid(*axes); return np.ascontiguousarray(np.concatenate([x.reshape(-1, 1) for x in xds], axis=1))
for (na, S, F) in [((10, 10, 10, 10), 4, 128), ((10,) * 6, 64, 1024)]:
    cand = mesh([np.linspace(0, 1, n) for n in na]); d = len(na)
    W, b, th = rs.randn(S, F, d) * 2, 2 * np.pi * rs.rand(S, F, 1), rs.randn(S, F, 1)
    det = ops.detect_product_structure(ops.dev(cand)); print("det", det, cand.shape, flush=True)
    res = {}
    for mode in (2, 3):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            i, v, fl = ops.rff_topk_product(cand, W, b, th, 1.3, det[0], det[1], det[2], k=3, tf32=mode)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res[mode] = (i.cpu().numpy(), v.cpu().numpy(), fl.cpu().numpy())
        print("mode", mode, "%.2f ms" % (dt * 1e3), "flags", int(fl.sum()), flush=True)
    print("idx equal", np.array_equal(res[2][0], res[3][0]), "val equal", np.array_equal(res[2][1], res[3][1]), flush=True)
