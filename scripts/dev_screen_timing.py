import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import ops
torch.cuda.set_device(0)
def run(N, d, S, F, k=1):
    g = torch.Generator(device="cuda").manual_seed(1)
    cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
    b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
    th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
    res = {}
    for m in (ops.RFF_FP64, ops.RFF_SCREEN, ops.RFF_SCREEN_MMA):
        ops.rff_topk(cand, W, b, th, 1.4, k, method=m); torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res[m] = ops.rff_topk(cand, W, b, th, 1.4, k, method=m); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        print("N=%d d=%d S=%d F=%d k=%d method=%d: %.4fs  %.3f Gpairs/s" % (N, d, S, F, k, m, best, N * S / best * 1e-9))
    print("  identical:", all(torch.equal(res[1][0], res[m][0]) and torch.equal(res[1][1], res[m][1]) for m in (2, 3)))
run(200_000, 6, 256, 1024)
run(1_000_000, 6, 1024, 1024)
run(1_000_000, 6, 1024, 1024, k=3)
run(1 << 20, 10, 64, 4096)
run(1 << 20, 2, 256, 512)
