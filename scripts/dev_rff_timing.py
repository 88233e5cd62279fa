"""Developer probe (not the bench): fp64 FMA peak and Stage-A timing at a few shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops

torch.cuda.set_device(0)
print("device", torch.cuda.get_device_name(0))
for s in (0.2, 1.0, 3.0):
    print("fp64 fma peak %.1fs: %.2f TFLOP/s" % (s, ops.fp64_fma_peak(s)))

def run(N, d, S, F, k=3, reps=2):
    g = torch.Generator(device="cuda").manual_seed(1)
    cand = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn(S, F, d, dtype=torch.float64, device="cuda", generator=g) * 1.5
    b = torch.rand(S, F, dtype=torch.float64, device="cuda", generator=g) * 6.283
    th = torch.randn(S, F, dtype=torch.float64, device="cuda", generator=g)
    ops.rff_topk(cand[: min(N, 4096)], W, b, th, 1.4, k)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.rff_topk(cand, W, b, th, 1.4, k); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    pairs = N * S
    dp_ops = pairs * F * (d + 13)
    print("N=%d d=%d S=%d F=%d: %.4f s  %.3f Gpairs/s  %.2f T dp-instr/s (x2 = %.2f TFLOP/s-equivalent)" % (
        N, d, S, F, best, pairs / best * 1e-9, dp_ops / best * 1e-12, 2 * dp_ops / best * 1e-12))

run(400, 2, 30, 300)
run(100_000, 6, 1024, 1024)
run(1_000_000, 6, 1024, 1024, reps=1)
run(1 << 20, 10, 64, 4096)
run(1 << 20, 2, 256, 512)
