"""Developer probe: TES_ep pipeline timing at C3-like shapes (random but well-conditioned inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tes_b200 import ops

torch.cuda.set_device(0)
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

def run(N, d, n, T, Sp):
    g = torch.Generator(device="cuda").manual_seed(1)
    l = torch.tensor([6.9512, 1.9341, 0.506, 4.2067, 5.0986, 3.5949][:d], dtype=torch.float64, device="cuda")
    x = torch.rand(N, d, dtype=torch.float64, device="cuda", generator=g)
    X = torch.rand(n, d, dtype=torch.float64, device="cuda", generator=g)
    Y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    tx = torch.rand(T, d, dtype=torch.float64, device="cuda", generator=g)
    sigma, sigma0 = 1.423, 1e-2
    pNK = ops.pnk(tx, X, l, sigma, sigma0)
    t_inv = timeit(lambda: ops.batched_chol2inv(pNK))
    inv = ops.batched_chol2inv(pNK)
    NK = ops.se_ard_cross(X, X, l, sigma) + torch.eye(n, dtype=torch.float64, device="cuda") * sigma0
    invK = ops.batched_chol2inv(NK)
    p = torch.rand(Sp, dtype=torch.float64, device="cuda", generator=g); p = p / p.sum()
    pm = torch.randn(Sp, T, dtype=torch.float64, device="cuda", generator=g)
    A = torch.randn(Sp, T, T, dtype=torch.float64, device="cuda", generator=g) * 0.05
    pc = A @ A.transpose(1, 2)
    t_stats = timeit(lambda: ops.posterior_stats(x, tx, X, Y, l, sigma, inv))
    t_det = timeit(lambda: ops.ep_acq(ops.EP_DETERMINISTIC, x, tx, X, Y, l, sigma, sigma0, inv, invK, p, pm, pc))
    t_lite = timeit(lambda: ops.ep_acq(ops.EP_LITE, x, tx, X, Y, l, sigma, sigma0, inv, invK, p, pm, pc))
    nb = (T + 15) // 16; kdim = nb * (nb + 1) // 2 * 256
    fl = 2.0 * N * Sp * kdim
    print("N=%d d=%d n=%d T=%d Sp=%d: pNK inv %.4fs | post-stats %.4fs | EP det %.4fs (quadform ~%.1f TFLOP/s if all) | lite %.4fs"
          % (N, d, n, T, Sp, t_inv, t_stats, t_det, fl / t_det * 1e-12, t_lite))

run(100_000, 6, 64, 32, 32)
run(1_000_000, 6, 64, 128, 128)
run(200_000, 6, 64, 256, 256)
# C5-style batched chol2inv
for m in (64, 128, 256, 512):
    batch = 1024 if m <= 256 else 256
    g = torch.Generator(device="cuda").manual_seed(m)
    Xp = torch.rand(batch, m, 6, dtype=torch.float64, device="cuda", generator=g)
    D = ((Xp[:, :, None, :] - Xp[:, None, :, :]) ** 2).sum(-1)
    A = 1.4 * torch.exp(-0.5 * 3.0 * D) + 1e-2 * torch.eye(m, dtype=torch.float64, device="cuda")
    t = timeit(lambda: ops.batched_chol2inv(A))
    fl = batch * (m ** 3 / 3.0 + m ** 3)
    print("chol2inv batch=%d m=%d: %.4fs  %.2f TFLOP/s (algorithmic 4/3 m^3)  %.1f GB/s (2 m^2 8B)" % (batch, m, t, fl / t * 1e-12, batch * 2 * m * m * 8 / t * 1e-9))
