"""Timing of the screen methods on one shape: python dev_tc5_time.py N d S F k [npoly list]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import ops
torch.cuda.set_device(0)
N, d, S, F, k = [int(a) for a in sys.argv[1:6]]
npolys = [a for a in sys.argv[6:]] or ["1"]   # "npoly" or "npoly:pc" (pc = polynomial columns of the specialised variant)
g = torch.Generator(device="cuda").manual_seed(1)
cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
ref = ops.rff_topk(cand, W, b, th, 1.4, k, method=3)
for m, npoly in [(3, None)] + [((5, n[1:]) if n.startswith('t') else (4, n)) for n in npolys]:
    if npoly is not None:
        np_, _, pc_ = str(npoly).partition(":")
        os.environ["TES_TC5_NPOLY"] = np_
        os.environ["TES_TC5_PC"] = pc_ or "0"
    r = ops.rff_topk(cand, W, b, th, 1.4, k, method=m); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ops.rff_topk(cand, W, b, th, 1.4, k, method=m); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    print("N=%d d=%d S=%d F=%d k=%d method %d npoly %s: %.4fs %.3f Gpairs/s %.2f Tcos/s identical=%s" % (
        N, d, S, F, k, m, npoly, best, N * S / best * 1e-9, N * S * F / best * 1e-12,
        torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1])), flush=True)
