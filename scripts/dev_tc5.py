"""Developer check of the tcgen05 screen (method 4): argument-error probe, identity with the fp64 kernel, timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops, _lib
from tes_b200.device import ptr, stream_ptr
torch.cuda.set_device(0)
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(1)
out = torch.zeros(1, dtype=torch.float64, device="cuda")
for d in (1, 2, 5, 6, 10):
    x = torch.rand(128, d, dtype=torch.float64, device="cuda", generator=g) * 10
    w = torch.randn(64 * 8, d, dtype=torch.float64, device="cuda", generator=g) * 3
    b = torch.rand(64 * 8, dtype=torch.float64, device="cuda", generator=g) * 6.283
    _lib.check(L.tes_tc5_arg_error_probe(ptr(x), ptr(w), ptr(b), d, 8, ptr(out), stream_ptr()), "probe")
    torch.cuda.synchronize()
    print("d=%d tc5 arg error / A = %.3e" % (d, float(out.cpu()[0])), flush=True)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
shapes = [(5000, 6, 16, 160, 3), (70_000, 2, 33, 300, 1), (40_000, 10, 64, 200, 8)] if small else \
         [(5000, 6, 16, 160, 3), (400_000, 6, 512, 1024, 1), (200_000, 10, 256, 1000, 3), (300_000, 2, 256, 512, 2)]
for (N, d, S, F, k) in shapes:
    cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
    b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
    th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
    ref = ops.rff_topk(cand, W, b, th, 1.4, k, method=ops.RFF_FP64)
    for m in [3, 4]:
        for npoly in ([None] if m == 3 else [0, 1, 2, 3, 4]):
            if npoly is not None:
                os.environ["TES_TC5_NPOLY"] = str(npoly)
            r = ops.rff_topk(cand, W, b, th, 1.4, k, method=m); torch.cuda.synchronize()
            best = 1e9
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = ops.rff_topk(cand, W, b, th, 1.4, k, method=m); e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            print("N=%d d=%d S=%d F=%d k=%d method %d npoly %s: %.4fs %.3f Gpairs/s %.2f Tcos/s identical=%s" % (
                N, d, S, F, k, m, npoly, best, N * S / best * 1e-9, N * S * F / best * 1e-12,
                torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1])), flush=True)
