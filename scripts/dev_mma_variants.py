import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import ops
torch.cuda.set_device(0)
N, d, S, F, k = 400_000, 6, 512, 1024, 1
g = torch.Generator(device="cuda").manual_seed(1)
cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
ref = ops.rff_topk(cand, W, b, th, 1.4, k, method=ops.RFF_FP64)
for v in ["4,0", "4,1", "4,2", "4,3", "4,4", "4,5"]:
    os.environ["TES_MMA_VARIANT"] = v
    ops.rff_topk(cand, W, b, th, 1.4, k, method=3); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ops.rff_topk(cand, W, b, th, 1.4, k, method=3); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    print("variant %s: %.4fs %.3f Gpairs/s identical=%s" % (v, best, N * S / best * 1e-9, torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1])))
