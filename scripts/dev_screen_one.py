import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import ops
torch.cuda.set_device(0)
N, d, S, F, k, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
g = torch.Generator(device="cuda").manual_seed(1)
cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
for _ in range(2):
    ops.rff_topk(cand, W, b, th, 1.4, k, method=m)
torch.cuda.synchronize()
print("ok")
