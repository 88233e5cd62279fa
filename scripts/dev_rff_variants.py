"""Developer probe: sweep rff_topk kernel variants (needs the library built with -DTES_RFF_DEV_VARIANTS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tes_b200 import ops

torch.cuda.set_device(0)
peak = ops.fp64_fma_peak(0.5)
print("fp64 peak %.2f" % peak)

def run(N, d, S, F, variants):
    g = torch.Generator(device="cuda").manual_seed(1)
    cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
    b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
    th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
    ref = None
    for v in variants:
        os.environ["TES_RFF_VARIANT"] = v
        try:
            idx, val = ops.rff_topk(cand, W, b, th, 1.4, 3)
        except Exception as e:
            print(v, "failed", e); continue
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx, val = ops.rff_topk(cand, W, b, th, 1.4, 3); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        if ref is None: ref = (idx.clone(), val.clone())
        ok = torch.equal(idx, ref[0]) and torch.equal(val, ref[1])
        tf = 2.0 * N * S * F * (d + 13) / best * 1e-12
        print("d=%d N=%d S=%d F=%d variant %-10s %.4fs  %.2f TF-eq  %.1f%% of peak  bit-exact=%s" % (d, N, S, F, v, best, tf, 100 * tf / peak, ok))

V = ["2,256,2", "2,256,1", "2,256,4", "4,256,1", "4,256,2", "2,512,2", "4,128,2", "3,256,2", "3,256,1", "4,512,1", "1,512,4", "1,256,4", "2,128,2"]
run(400_000, 6, 512, 1024, V)
run(400_000, 2, 512, 1024, V)
run(400_000, 10, 256, 1024, V)
