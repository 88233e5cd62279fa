"""Dev: sweeps / time of tes_sym_eig_f64 on C3-like Gram matrices Sigma_j = Z_j^T Z_j + sigma0 I (S=1024, n=64)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tes_b200 import ops

wl = bench.make_workload("c3", 0, 1)
h = wl["hyp"]
S, F, n, d = 256, wl["F"], wl["n"], wl["d"]
W = wl["randn_W"][:S] * np.sqrt(h["l"])
b = 2 * np.pi * wl["rand_b"][:S]
Z = np.sqrt(2 * h["sigma"] / F) * np.cos(W @ wl["X"].T[None] + b)     # (S,F,n)
Sig = np.transpose(Z, (0, 2, 1)) @ Z + h["sigma0"] * np.eye(n)
for nm, A in (("gram", Sig), ("random spd", None)):
    if A is None:
        rs = np.random.RandomState(0)
        M = rs.randn(S, n, n)
        A = M @ np.transpose(M, (0, 2, 1)) / n + 1e-3 * np.eye(n)
    At = torch.from_numpy(A).cuda()
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev, V, sw = ops.sym_eig(At, return_sweeps=True)
        e1.record(); torch.cuda.synchronize()
    sw = sw.cpu().numpy()
    ref = np.linalg.eigvalsh(A)
    got = np.sort(ev.cpu().numpy(), axis=1)
    print(nm, "ms %.3f" % e0.elapsed_time(e1), "sweeps min/mean/max", sw.min(), sw.mean(), sw.max(),
          "max rel eval err %.2e" % np.max(np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)),
          "cond %.1e" % (ref.max() / ref.min()))
A1 = torch.from_numpy(Sig[0]).cuda()
rs = np.random.RandomState(1)
M = rs.randn(128, 128); A128 = torch.from_numpy(M @ M.T / 128 + 1e-3 * np.eye(128)).cuda()
for nm, At in (("single 64", A1), ("single 128", A128)):
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev, V, sw = ops.sym_eig(At, return_sweeps=True)
        e1.record(); torch.cuda.synchronize()
    print(nm, "ms %.3f" % e0.elapsed_time(e1), "sweeps", int(sw[0]))
