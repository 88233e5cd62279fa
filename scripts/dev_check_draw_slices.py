import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from tes_b200 import optfunc
wl = bench.make_workload("c3", 0, 1)
dv = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("X", "Y", "randn_W", "rand_b", "noise")}
hyp = wl["hyp"]
full = optfunc.draw_random_init_weights_features(wl["d"], wl["S"], wl["F"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
                                                 draws=(dv["randn_W"], dv["rand_b"], dv["noise"]))
for world in (2, 8):
    per = wl["S"] // world
    ok = True
    for r in range(world):
        lo, hi = r * per, (r + 1) * per
        part = optfunc.draw_random_init_weights_features(wl["d"], per, wl["F"], dv["X"], dv["Y"], hyp["l"], hyp["sigma"], hyp["sigma0"],
                                                         draws=(dv["randn_W"][lo:hi], dv["rand_b"][lo:hi], dv["noise"][lo:hi]))
        for a, b in zip(full, part):
            ok &= bool(torch.equal(a[lo:hi], b))
    print("world", world, "slices bit-identical:", ok)
